// CUDA kernels for the batched r2r transforms (sm_100a).
//   fft_tile_kernel   : fast path, one tile of lines per CTA iteration, phases
//                       from fft_phases.cuh separated by __syncthreads()
//   r2r_direct_kernel : O(n^2) evaluation of the FFTW definitions for sizes /
//                       kinds the fast path does not cover (odd n, large prime
//                       factors, REDFT00/11, RODFT00/11); slow but complete.
#pragma once
#include <cuda_runtime.h>
#include "fft_phases.cuh"

namespace cb {

template <class T, class Lay> __device__ __forceinline__ Lay make_lay(const FftArgs<T>& A);
template <> __device__ __forceinline__ LayX make_lay<double, LayX>(const FftArgs<double>& A) { return LayX{LayX::line_len(A.P.M)}; }
template <> __device__ __forceinline__ LayX make_lay<float, LayX>(const FftArgs<float>& A) { return LayX{LayX::line_len(A.P.M)}; }
template <> __device__ __forceinline__ LayY make_lay<double, LayY>(const FftArgs<double>& A) { return LayY{A.tile_lines}; }
template <> __device__ __forceinline__ LayY make_lay<float, LayY>(const FftArgs<float>& A) { return LayY{A.tile_lines}; }

template <class T, class Lay>
__global__ void __launch_bounds__(256, 2) fft_tile_kernel(const FftArgs<T> A, long long ntiles) {
  extern __shared__ __align__(16) unsigned char cb_smem_raw[];
  T* s = reinterpret_cast<T*>(cb_smem_raw);
  const Lay lay = make_lay<T, Lay>(A);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool fwd = kind_is_forward(A.P.kind);
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const Tile tl = make_tile(A, t);
    if (fwd) phase_fwd_load(A, tl, s, lay, tid, nthr);
    else phase_bwd_pre(A, tl, s, lay, tid, nthr);
    __syncthreads();
    for (int st = 0; st < A.P.nstages; ++st) {
      phase_stage(A, tl, s, lay, st, tid, nthr);
      __syncthreads();
    }
    if (fwd) phase_fwd_post(A, tl, s, lay, tid, nthr);
    else phase_bwd_out(A, tl, s, lay, tid, nthr);
    __syncthreads();
  }
}

// ------------------------------------------------------------- direct O(n^2)
template <class T> struct DirectArgs {
  int n, kind, Q;            // table period is 2Q
  const C2<T>* cs;           // (cos, sin)(pi m / Q), m = 0..2Q-1
  const T* in;
  T* out;
  long long in_es, out_es, in_ls, out_ls, in_gs, out_gs;
  int lines_per_group, ngroups, line_len;
};

template <class T>
__global__ void __launch_bounds__(128) r2r_direct_kernel(const DirectArgs<T> A) {
  extern __shared__ __align__(16) unsigned char cb_smem_raw[];
  T* x = reinterpret_cast<T*>(cb_smem_raw);
  const long long line = blockIdx.x;
  const int g = (int)(line / A.lines_per_group);
  const int l = (int)(line - (long long)g * A.lines_per_group);
  const T* src = A.in + (long long)g * A.in_gs + (long long)l * A.in_ls;
  T* dst = A.out + (long long)g * A.out_gs + (long long)l * A.out_ls;
  const int n = A.n;
  for (int i = threadIdx.x; i < A.line_len; i += blockDim.x) x[i] = src[(long long)i * A.in_es];
  __syncthreads();
  const long long P2 = 2LL * A.Q;
  for (int k = threadIdx.x; k < A.line_len; k += blockDim.x) {
    if (k >= n) { dst[(long long)k * A.out_es] = x[k]; continue; }
    double acc = 0.0;
    // generic form: acc = extra + scale * sum_{j=j0}^{j1-1} x[j] * trig(pi * m_j / Q), m_j = m0 + j*step (mod 2Q)
    int j0 = 0, j1 = n, use_sin = 0;
    long long m0 = 0, step = 0;
    double scale = 2.0, extra = 0.0;
    switch (A.kind) {
      case K_R2HC:
        scale = 1.0;
        if (k <= n / 2) { step = (2LL * k) % P2; }
        else { step = (2LL * (n - k)) % P2; use_sin = 1; scale = -1.0; }
        break;
      case K_HC2R: break;  // handled below
      case K_REDFT00: j0 = 1; j1 = n - 1; step = k % P2; m0 = step; extra = (double)x[0] + ((k & 1) ? -1.0 : 1.0) * (double)x[n - 1]; break;
      case K_REDFT10: step = (2LL * k) % P2; m0 = k % P2; break;
      case K_REDFT01: j0 = 1; step = (2LL * k + 1) % P2; m0 = step; extra = (double)x[0]; break;
      case K_REDFT11: step = (2LL * (2LL * k + 1)) % P2; m0 = (2LL * k + 1) % P2; break;
      case K_RODFT00: use_sin = 1; step = (k + 1LL) % P2; m0 = step; break;
      case K_RODFT10: use_sin = 1; step = (2LL * (k + 1)) % P2; m0 = (k + 1LL) % P2; break;
      case K_RODFT01: use_sin = 1; j1 = n - 1; step = (2LL * k + 1) % P2; m0 = step; extra = ((k & 1) ? -1.0 : 1.0) * (double)x[n - 1]; break;
      case K_RODFT11: use_sin = 1; step = (2LL * (2LL * k + 1)) % P2; m0 = (2LL * k + 1) % P2; break;
      default: break;
    }
    if (A.kind == K_HC2R) {
      // x_k = r0 + 2 sum_f (r_f cos(2 pi f k/n) - i_f sin(2 pi f k/n)) + (-1)^k r_{n/2}
      acc = (double)x[0];
      const long long st2 = (2LL * k) % P2;
      long long m = st2;
      for (int f = 1; f < (n + 1) / 2; ++f) {
        const C2<T> w = A.cs[m];
        acc += 2.0 * ((double)x[f] * (double)w.x - (double)x[n - f] * (double)w.y);
        m += st2; if (m >= P2) m -= P2;
      }
      if ((n & 1) == 0) acc += ((k & 1) ? -1.0 : 1.0) * (double)x[n / 2];
    } else {
      long long m = m0;
      // m0 above is the phase of j = j0 for kinds whose first term has j = j0:
      //   forms (2j+1)*g -> m(j=0) = g, step 2g ; forms (j+1)*g or j*g with j0=1 -> m(j0) = g, step g
      double sum = 0.0;
      for (int j = j0; j < j1; ++j) {
        const C2<T> w = A.cs[m];
        sum += (double)x[j] * (double)(use_sin ? w.y : w.x);
        m += step; if (m >= P2) m -= P2;
      }
      acc = extra + scale * sum;
    }
    dst[(long long)k * A.out_es] = (T)acc;
  }
}

}  // namespace cb
