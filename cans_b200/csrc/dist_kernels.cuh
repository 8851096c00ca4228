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

// Flag words: every rank owns CB_FLAG_SLOTS x CB_MAX_RANKS words at the start of its exchange region; word
// [slot][r] is written by rank r only.  The value is the sequence number of the distributed solve (one per
// cansb200_solve / _solve_z call, the same on every rank because the calls are collective), so flags never
// need a reset and a rank that fails half-way through a solve cannot shift the meaning of later flags.
// Slots: 0..7 "my forward y transform of x window w has landed on every z pencil", 8..15 "my tridiagonal
// solve of x window w has landed on every slab", 16 / 17 whole-field barriers (solve_z, distributed TDMA).
#define CB_FLAG_SLOTS 32
#define CB_SLOT_FWD 0
#define CB_SLOT_BWD 8
#define CB_SLOT_BAR 16
#define CB_MAX_WINDOWS 8

// One CTA, thread s talks to rank s.  SIGNAL: announce "everything enqueued on this stream before this kernel
// is done" (kernel boundary + system fence order the earlier peer stores before the flag).  WAIT: spin until
// every rank has announced the same.  A rank that never shows up trips the timeout instead of hanging the GPU;
// `status` is a host-mapped word the library checks before the next solve (CANSB200_ECOMM), `status_dev` its
// device-resident twin (what later waits look at).
// `target` >= 0: signal that one rank only (the copy-engine exchange announces every transfer to its receiver);
// `skip_self`: the wait leaves out this rank's own word (nobody announces a transfer to itself).
__global__ void dist_flag_kernel(DistPeers peers, int rank, int nranks, int slot, unsigned long long seq, int do_signal,
                                 int do_wait, volatile int* status, volatile int* status_dev, unsigned long long timeout_ns,
                                 int target, int skip_self) {
  const int s = threadIdx.x;
  if (s >= nranks) return;
  if (do_signal && (target < 0 || target == s)) {
    __threadfence_system();
    st_release_sys(peers.flags[s] + slot * CB_MAX_RANKS + rank, seq);
  }
  if (!do_wait || *status_dev || (skip_self && s == rank)) return;   // after one time-out the solve is lost anyway: later waits do not add 20 s each
  const unsigned long long* mine = peers.flags[rank] + slot * CB_MAX_RANKS + s;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(mine) < seq) {
    if (global_timer_ns() - t0 > timeout_ns) {
      *status_dev = 1;
      *status = 1;
      __threadfence_system();
      break;
    }
    __nanosleep(100);
  }
  __threadfence_system();
}

// rows [k0, k1) of a (rows x ncol) array to their peer-mapped homes (used for rows the tridiagonal
// kernel does not write: the face-centred Dirichlet plane, and the sequential fallback's result);
// rows[k] = {ptr, sj}: element (j, i) of row k goes to ptr[j * sj + i]
template <class T> struct DistOutRow { T* ptr; long long sj; };
template <class T>
__global__ void scatter_rows_kernel(const T* __restrict__ src, long long sk, const DistOutRow<T>* __restrict__ rows, int k0, int k1,
                                    int ny, int nx, int xb, int xn) {
  const long long ncolw = (long long)ny * xn;   // columns of the window [xb, xb + xn) of every j
  const long long tot = (long long)(k1 - k0) * ncolw;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const long long k = k0 + e / ncolw, cw = e - (e / ncolw) * ncolw;
    const long long j = cw / xn, i = xb + (cw - j * xn);
    rows[k].ptr[j * rows[k].sj + i] = src[k * sk + j * nx + i];
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
