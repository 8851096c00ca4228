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

// ---- fillps and correc with a 3-D launch geometry: x along threadIdx.x (coalesced), y along threadIdx.y, one z plane per
// blockIdx.z.  No index division per element (the flat kernels above spend two 64-bit divisions per point); same expressions,
// same operation order.  Used whenever the extents fit the grid limits.  Measured on C3 (profiles/r2_bench_steps.json):
// fillps 1.84 -> 1.34 ms (6.4 TB/s), correc 3.33 -> 2.89 ms (5.2 TB/s).  chkdiv keeps its capped grid-stride kernel: with one
// CTA per 256 points the two atomics per CTA made it slower (3.9 against 2.8 ms), so that variant is not kept.
// launch geometry of these kernels: 256 threads, x first; false when an extent exceeds the grid limits (flat kernels then)
inline bool aux_geom(int ex, int ey, int ez, dim3& grid, dim3& block) {
  int bx = 32;
  while (bx < 256 && bx < ex) bx *= 2;
  const int by = 256 / bx;
  const long long gy = ((long long)ey + by - 1) / by;
  if (ex < 1 || ey < 1 || ez < 1 || gy > 65535 || ez > 65535) return false;
  block = dim3((unsigned)bx, (unsigned)by, 1);
  grid = dim3((unsigned)((ex + bx - 1) / bx), (unsigned)gy, (unsigned)ez);
  return true;
}

template <class T>
__global__ void fillps3d_kernel(int n1, int n2, T dxi, T dyi, const T* __restrict__ dzfi, T dti, const T* __restrict__ u,
                                const T* __restrict__ v, const T* __restrict__ w, T* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i >= n1 || j >= n2) return;
  const long long p1 = n1 + 2, p2 = n2 + 2;
  const T dtidxi = dti * dxi, dtidyi = dti * dyi;
  const long long o = ((long long)(k + 1) * p2 + (j + 1)) * p1 + (i + 1);
  p[o] = (w[o] - w[o - p1 * p2]) * dti * dzfi[k + 1] + (v[o] - v[o - p1]) * dtidyi + (u[o] - u[o - 1]) * dtidxi;
}

template <class T>
__global__ void correc3d_kernel(int n1, int n2, int n3, T dxi, T dyi, const T* __restrict__ dzci, T dt, const T* __restrict__ p,
                                T* __restrict__ u, T* __restrict__ v, T* __restrict__ w) {
  // haloed extents, as the reference's three loops (src/correc.f90:33-59): u for i <= n1, v for j <= n2, w for k <= n3
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i > n1 + 1 || j > n2 + 1) return;
  const long long p1 = n1 + 2, p2 = n2 + 2;
  const long long e = ((long long)k * p2 + j) * p1 + i;
  const T fi = dt * dxi, fj = dt * dyi;
  const T pc = p[e];
  if (i <= n1) u[e] = u[e] - fi * (p[e + 1] - pc);
  if (j <= n2) v[e] = v[e] - fj * (p[e + p1] - pc);
  if (k <= n3) w[e] = w[e] - dt * dzci[k] * (p[e + p1 * p2] - pc);
}

// ---- eigenvalue order of an _OPENACC-built `initsolver` ------------------------------------------------------------
// The CUDA build of the reference stores the spectrum of a periodic direction as cuFFT leaves it after its own
// post-processing, (r0, r[n/2], r1, i1, r2, i2, ...), and permutes the eigenvalues to match
// (/root/reference/src/initsolver.f90:98-117, `iswap`).  The kernels here keep FFTW's halfcomplex order
// (r0, r1, ..., r[n/2], i[n/2-1], ..., i1), which is what the CPU build's `initsolver` produces.  pack_index maps a
// halfcomplex position h to the position of the same mode in the packed order, so that
//     lambda_halfcomplex[h] = lambda_packed[pack_index(h, n)].
__host__ __device__ inline int pack_index(int h, int n) {
  if (h == 0) return 0;
  if (2 * h <= n) return (2 * h == n) ? 1 : 2 * h;   // real part of mode h (r[n/2] sits at 1 when n is even)
  const int k = n - h;                                 // imaginary part of mode k
  return (2 * k + 1 < n) ? 2 * k + 1 : 1;              // odd n: i[(n-1)/2] takes the slot r[n/2] has for even n
}
// halfcomplex position of position `pos` of the SPLIT order (r0 .. r[n/2-1] | r[n/2], i1 .. i[n/2-1]; CB_R2_XSPLIT)
__host__ __device__ inline int split_to_hc(int pos, int n) { return 2 * pos <= n ? pos : n - (pos - n / 2); }
// out[j][i] = in[qy(j)][qx(i)]: the eigenvalues in the order the kernels keep the spectrum in (x: split order when sx, else
// halfcomplex; y: halfcomplex), gathered from the caller's order (px / py: packed in that direction, else halfcomplex)
template <class T>
__global__ void lambda_unpack_kernel(const T* __restrict__ in, T* __restrict__ out, int nx, int ny, int px, int py, int sx) {
  const long long tot = (long long)nx * ny;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e / nx), i = (int)(e - (long long)j * nx);
    int qi = sx ? split_to_hc(i, nx) : i;
    qi = px ? pack_index(qi, nx) : qi;
    const int qj = py ? pack_index(j, ny) : j;
    out[e] = in[(long long)qj * nx + qi];
  }
}

// updt_rhs_b (src/bound.f90:514-598): p(plane) += value * norm on up to six boundary planes of the interior.
// plane[d][s] = interior index (1-based) of the plane of direction d, side s, or 0 when this rank does not own that wall
// or no value was passed.
struct RhsbPlanes { int idx[3][2]; double val[3][2]; };
template <class T>
__global__ void updt_rhs_b_kernel(T* __restrict__ p, int n1, int n2, int n3, RhsbPlanes B, int dir) {
  // one launch per direction (edge points belong to planes of two directions: the reference adds them one loop after
  // the other, so must we); one thread per face point adds the lower then the upper plane (they coincide when n = 1)
  const long long p1 = n1 + 2, p2 = n2 + 2;
  const long long tot = dir == 0 ? (long long)n2 * n3 : (dir == 1 ? (long long)n1 * n3 : (long long)n1 * n2);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int q = B.idx[dir][s];
      if (!q) continue;
      const T v = (T)B.val[dir][s];
      if (dir == 0) { const long long k = e / n2 + 1, j = e % n2 + 1; p[(k * p2 + j) * p1 + q] += v; }
      else if (dir == 1) { const long long k = e / n1 + 1, i = e % n1 + 1; p[(k * p2 + q) * p1 + i] += v; }
      else { const long long j = e / n1 + 1, i = e % n1 + 1; p[((long long)q * p2 + j) * p1 + i] += v; }
    }
  }
}

}  // namespace cb
