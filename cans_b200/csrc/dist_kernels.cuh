// Multi-GPU plumbing of the z-slab decomposition: the x and y transforms are local to a
// slab (nx, ny, nz/P); the tridiagonal stage needs (nx, ny/P, nz).  The reference moves the data with
// pack -> all-to-all -> unpack around every transpose (cuDecomp: /root/reference/dependencies/cuDecomp/
// include/internal/transpose.h:196-905; 2DECOMP: dependencies/2decomp-fft/src/transpose_y_to_z.f90).
// Here the producer kernels store their 128-byte rows straight into the consumer GPU's buffer through
// peer-mapped pointers (CUDA IPC over NVLink): pack, wire transfer and unpack are one store.  What is
// left is a barrier between the producers on all GPUs and the consumer, implemented on the device with
// system-scope release/acquire flags so that the solve stays stream ordered (no host synchronisation).
#pragma once
#include <cuda_runtime.h>

namespace cb {

#define CB_MAX_RANKS 16

struct DistPeers {
  unsigned long long* flags[CB_MAX_RANKS];   // flags[s] = flag array of rank s (peer mapped); slot [r] is written by rank r
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One CTA, thread s talks to rank s: announce "everything I enqueued before this kernel is done"
// (kernel boundary + system fence order the earlier peer stores before the flag), then wait for the same
// announcement of every rank.  A rank that never shows up trips the timeout instead of hanging the GPU.
__global__ void dist_barrier_kernel(DistPeers peers, int rank, int nranks, unsigned long long epoch, int* status,
                                    unsigned long long timeout_ns) {
  const int s = threadIdx.x;
  if (s >= nranks) return;
  __threadfence_system();
  st_release_sys(peers.flags[s] + rank, epoch);
  const unsigned long long* mine = peers.flags[rank] + s;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(mine) < epoch) {
    if (global_timer_ns() - t0 > timeout_ns) {
      atomicExch(status, 1);
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

// rows [k0, k1) of a (rows x ncol) array to their peer-mapped homes (used for rows the tridiagonal
// kernel does not write: the face-centred Dirichlet plane, and the sequential fallback's result);
// rows[k] = {ptr, sj}: element (j, i) of row k goes to ptr[j * sj + i]
template <class T> struct DistOutRow { T* ptr; long long sj; };
template <class T>
__global__ void scatter_rows_kernel(const T* __restrict__ src, long long sk, const DistOutRow<T>* __restrict__ rows, int k0, int k1,
                                    long long ncol, int nx) {
  const long long tot = (long long)(k1 - k0) * ncol;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long k = k0 + e / ncol, col = e - (e / ncol) * ncol;
    const long long j = col / nx, i = col - j * nx;
    rows[k].ptr[j * rows[k].sj + i] = src[k * sk + col];
  }
}

// z-only solve (solver_gaussel_z) on a decomposed grid: the x-pencil slab goes to the z pencils of its owners
// and back without any transform in between.  Row (j, g) of my haloed slab <-> tab[j].ptr + g * tab[j].gs
// (the same row tables the y transforms use for their peer-mapped stores / loads).
template <class T> struct DistRow { T* ptr; long long gs; };
template <class T>
__global__ void slab_rows_copy_kernel(T* __restrict__ slab, long long px, long long pxy, const DistRow<T>* __restrict__ tab,
                                      int nx, int ny, int nzl, int to_peers) {
  const long long tot = (long long)nx * ny * nzl;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % nx);
    const long long r = e / nx;
    const int j = (int)(r % ny), g = (int)(r / ny);
    T* mine = slab + (long long)g * pxy + (long long)j * px + i;
    T* theirs = tab[j].ptr + (long long)g * tab[j].gs + i;
    if (to_peers) *theirs = *mine;
    else *mine = *theirs;
  }
}

}  // namespace cb
