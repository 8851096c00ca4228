// Kernels for the steps either side of the solver path and for synthetic
// input: fillps (src/fillps.f90:38-50), correc (src/correc.f90:33-59), chkdiv
// (src/chkdiv.f90:35-50) and the counter-based hash field of SURVEY.md 8(d).
// Plain streaming kernels (one pass each); haloed arrays p(0:n1+1,0:n2+1,0:n3+1).
#pragma once
#include <cuda_runtime.h>

namespace cb {

__device__ __forceinline__ double hash_uniform(unsigned long long idx, unsigned long long seed) {
  unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ULL + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  const double u01 = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  return 2.0 * u01 - 1.0;
}

// twin of oracle.cans_oracle.hash_field
template <class T>
__global__ void fill_hash_kernel(T* p, int n1, int n2, int n3, int o1, int o2, int o3, int ng1, int ng2, int nh,
                                 unsigned long long seed) {
  const long long p1 = n1 + 2 * nh, p2 = n2 + 2 * nh, p3 = n3 + 2 * nh;
  const long long tot = p1 * p2 * p3;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / (p1 * p2), r = e - k * p1 * p2, j = r / p1, i = r - j * p1;
    const long long ii = i - nh, jj = j - nh, kk = k - nh;
    T v = T(0);
    if (ii >= 0 && ii < n1 && jj >= 0 && jj < n2 && kk >= 0 && kk < n3) {
      const unsigned long long gi = (unsigned long long)(ii + o1), gj = (unsigned long long)(jj + o2), gk = (unsigned long long)(kk + o3);
      v = (T)hash_uniform((gk * (unsigned long long)ng2 + gj) * (unsigned long long)ng1 + gi, seed);
    }
    p[e] = v;
  }
}

template <class T>
__global__ void fillps_kernel(int n1, int n2, int n3, T dxi, T dyi, const T* __restrict__ dzfi, T dti, const T* __restrict__ u,
                              const T* __restrict__ v, const T* __restrict__ w, T* __restrict__ p) {
  const long long p1 = n1 + 2, p2 = n2 + 2;
  const long long tot = (long long)n1 * n2 * n3;
  const T dtidxi = dti * dxi, dtidyi = dti * dyi;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / ((long long)n1 * n2), r = e - k * n1 * n2, j = r / n1, i = r - j * n1;
    const long long o = ((k + 1) * p2 + (j + 1)) * p1 + (i + 1);
    p[o] = (w[o] - w[o - p1 * p2]) * dti * dzfi[k + 1] + (v[o] - v[o - p1]) * dtidyi + (u[o] - u[o - 1]) * dtidxi;
  }
}

template <class T>
__global__ void correc_kernel(int n1, int n2, int n3, T dxi, T dyi, const T* __restrict__ dzci, T dt, const T* __restrict__ p,
                              T* __restrict__ u, T* __restrict__ v, T* __restrict__ w) {
  const long long p1 = n1 + 2, p2 = n2 + 2, p3 = n3 + 2;
  const long long tot = p1 * p2 * p3;
  const T fi = dt * dxi, fj = dt * dyi;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / (p1 * p2), r = e - k * p1 * p2, j = r / p1, i = r - j * p1;
    const T pc = p[e];
    if (i <= n1) u[e] = u[e] - fi * (p[e + 1] - pc);
    if (j <= n2) v[e] = v[e] - fj * (p[e + p1] - pc);
    if (k <= n3) w[e] = w[e] - dt * dzci[k] * (p[e + p1 * p2] - pc);
  }
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double val) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  const unsigned long long nv = (unsigned long long)__double_as_longlong(val);
  while (nv > old) {
    assumed = old;
    old = atomicCAS(a, assumed, nv);
    if (old == assumed) break;
  }
}

// res[0] += sum |div| * cell volume, res[1] = max |div|
template <class T>
__global__ void chkdiv_kernel(int n1, int n2, int n3, T dxi, T dyi, const T* __restrict__ dzfi, const T* __restrict__ u,
                              const T* __restrict__ v, const T* __restrict__ w, double* res) {
  const long long p1 = n1 + 2, p2 = n2 + 2;
  const long long tot = (long long)n1 * n2 * n3;
  double sum = 0.0, mx = 0.0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / ((long long)n1 * n2), r = e - k * n1 * n2, j = r / n1, i = r - j * n1;
    const long long o = ((k + 1) * p2 + (j + 1)) * p1 + (i + 1);
    const T div = (w[o] - w[o - p1 * p2]) * dzfi[k + 1] + (v[o] - v[o - p1]) * dyi + (u[o] - u[o - 1]) * dxi;
    const double ad = fabs((double)div);
    mx = fmax(mx, ad);
    sum += (double)(fabs(div) / (dxi * dyi * dzfi[k + 1]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&res[0], sum);
    atomic_max_nonneg(&res[1], mx);
  }
}

}  // namespace cb
