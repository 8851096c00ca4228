// Tridiagonal solve in z -- replaces `gaussel` / `gaussel_gpu`
// (/root/reference/src/solver.f90:114-307, src/solver_gpu.f90:278-428).
//
// Numerics (DESIGN.md "Thomas"): the reference's own result carries ~1e-11
// relative rounding error on nz = 512, ALL of it from the pivot recurrence
//   z_k = 1/((b_k + lambda) - a_k d_{k-1}),  d_k = c_k z_k
// while the two linear substitution sweeps contribute ~1e-15.  To stay within
// 1e-12 of the reference the pivots must be bit-identical to its sequence, so:
//   thomas_factor_kernel : one thread per column, exactly the reference's
//       operation order (no FMA contraction, IEEE division), including the
//       singular-pivot pin (:151-164) encoded as z = 0, and for periodic z the
//       auxiliary solve p2 (:201-260) and closure denominator (:276-283).
//       Runs only when (a, b, c, lambda) change: results are cached per plan,
//       keyed by a device-side content hash (no host sync).
//   thomas_seq_kernel    : one thread per column, two sweeps through global
//       memory using the cached pivots, reference operation order.  Any nz.
//   thomas_pipe_kernel   : persistent tiles of {one 128-byte row segment of
//       columns x all rows}, fetched by TMA while the previous tile is solved;
//       every thread keeps a chunk of m rows (values and pivots) in registers
//       and the two first-order linear recurrences are evaluated chunk-parallel
//       (local sweep, warp-shuffle scan of the affine chunk maps, fix-up).  p'
//       never leaves the chip: traffic is read p + read z + write p = 24 B/pt
//       instead of 48.  Above 512 rows a cluster of two CTAs shares the tile
//       and hands the carries over through distributed shared memory.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

// where result row k of the distributed solve lives: element (j, i) of my z pencil goes to ptr[j * sj + i] on the
// GPU that owns plane k.  The way-back buffer is ordered [j][k][i] per source rank, so that the rows one tile
// writes to a peer are 16 KB apart (dense 2 MB windows over NVLink), not a whole pencil plane apart.
template <class T> struct __align__(16) OutRow { T* ptr; long long sj; };

template <class T> struct ThomasDev {
  int nx, ny;         // columns (x fastest)
  int n;              // rows handled: nz - q, incl. the periodic closure row
  int nn;             // n-1 if periodic else n
  int periodic;
  long long sj, sk;   // field strides: p[k*sk + j*sj + i]
  const T* a; const T* b; const T* c;  // device, length >= n
  const T* lam; long long lam_sj;      // lambdaxy[j*lam_sj + i]
  const OutRow<T>* out_rows; // distributed solve: row k of the result goes to out_rows[k] (a peer-mapped pointer);
                      // nullptr = in place
  int xb, xn;         // column window of this launch: i in [xb, xb + xn) for every j (thomas_reg_kernel)
  int m;              // rows per chunk of the chunked substitution = ceil(nn/32)
  int chunk_layout;   // 2 = z[j][k][i] (the only layout, see zidx); kept for ABI stability of the struct
  int nopin;          // 1: the lambda-less variant of gaussel (src/solver.f90:168-188, :238-256; solver_gaussel_z):
                      // no singular-pivot pin, no tolerance test on the periodic closure
  // Pivot-cache deduplication.  The pivots of column (i, j) depend on lambdaxy(i, j) only, and in a periodic direction the
  // real and the imaginary part of a mode share their eigenvalue.
  //   y (dy): rows are in halfcomplex order, lambda(j) = lambda(ny - j): the cache keeps rows j <= ny/2, row j > ny/2 uses ny - j.
  //   x (dx): the solve keeps x in SPLIT order between the two x transforms (r0 .. r[n/2-1] | r[n/2], i1 .. i[n/2-1], see
  //           CB_R2_XSPLIT in r2r2.cuh), so that position nx/2 + p holds the imaginary part of the mode whose real part sits
  //           at p: whole 128-byte tiles pair up, in the same column order.  The cache keeps positions 0 .. nxu - 1 with
  //           nxu = nx/2 + one tile (the tile at nx/2 starts with r[n/2] and is stored on its own); position i >= nxu uses
  //           i - nx/2.
  // A quarter of the cache (and of its HBM stream) for a doubly periodic operator.  nxu / nyu = stored extents.
  int dx, dy, nxu, nyu;
  long long zsj, zsk;  // element strides of the pivot array between y rows and between z rows (cache: nn * nxu and nxu)
  // Distributed-TDMA mode of the pipelined kernel (dt_mode = 1): the slab-local elimination of gaussel_dtdma
  // (src/solver.f90:351-391) is the same pair of first-order recurrences as gaussel's two sweeps with
  //   a_1 := 0 (rows 0 and 1 both start a chain), c_{n-2} := 0 (rows n-2 and n-1 keep their forward value),
  //   row 0 scaled by Z1 = 1 / (1 - aa_1 cc_0) at the end,
  // pivots Z[k][j][i] from the coefficient cache of the plan; rows 0 and n-1 also go to the reduced right-hand side dt_rp.
  // Shallow grids: jb > 1 solves jb consecutive y rows as ONE tall tile of jb * nn rows (the z-major field and the pivot cache
  // are contiguous in (j, k), so the tile is still a uniform run of rows); the recurrences are cut at the seams by a_0 := 0
  // and c_{nn-1} := 0.  Keeps 512 rows per tile when nz is small (nz = 128: 2.8 -> ~1 ms on 2048 x 1024 x 128).
  int jb;
  int dt_mode;
  const T* dt_z1;      // [ncol] of the selected slot
  T* dt_rp;            // [2][ncol]
  long long dt_slot_small;   // elements between the slots of dt_z1
};

// ---- exactly-rounded, never-contracted arithmetic ---------------------------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
template <class T> __device__ __forceinline__ T eps_of();
template <> __device__ __forceinline__ double eps_of<double>() { return 2.220446049250313e-16; }
template <> __device__ __forceinline__ float eps_of<float>() { return 1.1920929e-07f; }

// Pivot-cache layout: z[j][k][i] (row pitch nx), so that the nn rows of one column tile sit in one window of
// nn * nx elements (a 4 MB window on C3) instead of one row per field plane: the tile fetch then has the access
// pattern of the y transforms for half of its bytes.
template <class T> __device__ __forceinline__ int zux(const ThomasDev<T>& D, int i) { return (D.dx && i >= D.nxu) ? i - D.nx / 2 : i; }
template <class T> __device__ __forceinline__ int zuy(const ThomasDev<T>& D, int j) { return (D.dy && 2 * j > D.ny) ? D.ny - j : j; }
template <class T> __device__ __forceinline__ long long zidx(const ThomasDev<T>& D, int i, int j, int k) {
  return ((long long)zuy(D, j) * D.nn + k) * D.nxu + zux(D, i);
}
template <class T> __device__ __forceinline__ long long zden(const ThomasDev<T>& D, int i, int j) {
  return (long long)zuy(D, j) * D.nxu + zux(D, i);
}

// ---- factorisation cache bookkeeping (device resident) ----------------------
#define CB_MAX_SLOTS 8
struct CacheState {
  unsigned long long key_new, key2_new;   // two independent content hashes of the coefficient set being solved
  unsigned long long keys[CB_MAX_SLOTS];
  unsigned long long keys2[CB_MAX_SLOTS];
  unsigned long long stamp[CB_MAX_SLOTS];
  unsigned long long clock;
  unsigned long long nfactor;  // how many factorisations ran (diagnostic)
  int sel, hit, nslots;
  int sym_bad;                 // deduplicated cache: lambdaxy was found NOT mirror-symmetric (sticky; the host falls back / reports)
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// second, independent mixer (murmur3 finaliser with a different stream constant): a slot is reused only if BOTH 64-bit
// hashes match, so a silent reuse of the wrong factorisation needs a simultaneous collision of two unrelated sums
__device__ __forceinline__ unsigned long long mix64b(unsigned long long z) {
  z ^= 0xC2B2AE3D27D4EB4FULL;
  z = (z ^ (z >> 33)) * 0xFF51AFD7ED558CCDULL;
  z = (z ^ (z >> 33)) * 0xC4CEB9FE1A85EC53ULL;
  return z ^ (z >> 33);
}
__device__ __forceinline__ unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long bits_of(float v) { return (unsigned long long)__float_as_uint(v); }

// order-independent content hash of (a, b, c, lambdaxy)
template <class T>
__global__ void thomas_hash_kernel(const ThomasDev<T> D, CacheState* st) {
  const long long ncoef = D.n;
  const long long nlam = (long long)D.nx * D.ny;
  const long long total = 3 * ncoef + nlam;
  unsigned long long h = 0, h2 = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    T v;
    if (e < ncoef) v = D.a[e];
    else if (e < 2 * ncoef) v = D.b[e - ncoef];
    else if (e < 3 * ncoef) v = D.c[e - 2 * ncoef];
    else {
      const long long q = e - 3 * ncoef;
      const long long j = q / D.nx, i = q - j * D.nx;
      v = D.lam[j * D.lam_sj + i];
      if (D.dx | D.dy) {
        // The deduplicated cache is only valid for a mirror-symmetric lambdaxy: check every element, every solve.  initsolver's
        // eigenvalues are symmetric up to the rounding of cos(2 pi (n - l) / n) against cos(2 pi l / n): <= 2e-13 relative for
        // n = 1024 (2e-11 for n = 2048, lowest modes); replacing lambda(n - l) by lambda(l) moves the solution of such a column
        // by a third of that, i.e. <= 1e-14 of the field in relative L2 (measured against the oracle: tests).
        const T u = D.lam[(long long)zuy(D, (int)j) * D.lam_sj + zux(D, (int)i)];
        const T big = fabs(u) > fabs(v) ? fabs(u) : fabs(v);
        if (fabs(u - v) > T(1e-10) * big) st->sym_bad = 1;
      }
    }
    h += mix64(mix64((unsigned long long)e) ^ bits_of(v));
    h2 += mix64b(mix64b(bits_of(v)) + 0x9E3779B97F4A7C15ULL * (unsigned long long)(e + 1));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) h += mix64(0x5EEDULL + (unsigned long long)D.nopin + 2ULL * D.dx + 4ULL * D.dy);
  for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(0xffffffffu, h, o); h2 += __shfl_xor_sync(0xffffffffu, h2, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&st->key_new, h); atomicAdd(&st->key2_new, h2); }
}

__global__ void thomas_select_kernel(CacheState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long key = st->key_new | 1ULL;  // never 0 (0 = empty slot)
  const unsigned long long key2 = st->key2_new;
  st->clock += 1;
  int sel = -1;
  for (int s = 0; s < st->nslots; ++s)
    if (st->keys[s] == key && st->keys2[s] == key2) sel = s;
  if (sel >= 0) {
    st->hit = 1;
  } else {
    st->hit = 0;
    sel = 0;
    for (int s = 1; s < st->nslots; ++s)
      if (st->stamp[s] < st->stamp[sel]) sel = s;
    st->keys[sel] = key;
    st->keys2[sel] = key2;
    st->nfactor += 1;
  }
  st->stamp[sel] = st->clock;
  st->sel = sel;
  st->key_new = 0;
  st->key2_new = 0;
}

// ---- factorisation: reference operation order, one thread per column --------
template <class T>
__global__ void __launch_bounds__(128) thomas_factor_kernel(const ThomasDev<T> D, const CacheState* st, T* zbase, T* p2base,
                                                            T* denbase, long long slot_z, long long slot_den) {
  if (st->hit) return;
  const int sel = st->sel;
  T* z = zbase + (long long)sel * slot_z;
  T* p2 = p2base ? p2base + (long long)sel * slot_z : nullptr;
  T* den_c = denbase ? denbase + (long long)sel * slot_den : nullptr;
  // one thread per STORED column: all of them, or positions i < nxu (rows j <= ny/2) of a deduplicated direction
  const int nxs = D.dx ? D.nxu : D.nx, nys = D.dy ? D.ny / 2 + 1 : D.ny;
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)nxs * nys) return;
  const int j = (int)(col / nxs), i = (int)(col - (long long)j * nxs);
  const T lam = D.lam[(long long)j * D.lam_sj + i];
  const int nn = D.nn;
  const T one = T(1);
  // first sweep: pivots with the singular-pivot pin, src/solver.f90:138-166
  T zz = div_rn(one, add_rn(D.b[0], lam));
  T d = mul_rn(D.c[0], zz);
  z[zidx(D, i, j, 0)] = zz;
  for (int k = 1; k < nn; ++k) {
    const T bl = add_rn(D.b[k], lam);
    const T ad = mul_rn(D.a[k], d);
    const T den = sub_rn(bl, ad);
    bool pin = false;
    if (k == nn - 1 && !D.nopin) {
      const T tol = mul_rn(eps_of<T>(), fmax(fabs(bl), fabs(ad)));
      pin = fabs(den) <= tol;
    }
    zz = pin ? T(0) : div_rn(one, den);
    d = mul_rn(D.c[k], zz);
    z[zidx(D, i, j, k)] = zz;
  }
  if (!D.periodic) return;
  // auxiliary system of the periodic closure, src/solver.f90:201-260 (no pin in this sweep)
  zz = div_rn(one, add_rn(D.b[0], lam));
  d = mul_rn(D.c[0], zz);
  T pv = (nn == 1) ? sub_rn(-D.a[0], D.c[0]) : -D.a[0];
  pv = mul_rn(pv, zz);
  p2[zidx(D, i, j, 0)] = pv;
  for (int k = 1; k < nn; ++k) {
    zz = div_rn(one, sub_rn(add_rn(D.b[k], lam), mul_rn(D.a[k], d)));
    d = mul_rn(D.c[k], zz);
    const T rhs = (k == nn - 1) ? sub_rn(T(0), D.c[k]) : T(0);
    pv = mul_rn(sub_rn(rhs, mul_rn(D.a[k], pv)), zz);
    p2[zidx(D, i, j, k)] = pv;
  }
  // back substitution needs d_k = c_k z_k of THIS sweep; recompute it forward is not possible
  // in reverse, so redo: d_k of this sweep equals c_k * z_k(unpinned); the pinned row (k = nn-1)
  // is never used as d in the back substitution (:261-270 runs k = nn-1..1, 1-based).
  for (int k = nn - 2; k >= 0; --k) {
    const T dk = mul_rn(D.c[k], z[zidx(D, i, j, k)]);
    pv = sub_rn(p2[zidx(D, i, j, k)], mul_rn(dk, pv));
    p2[zidx(D, i, j, k)] = pv;
  }
  const T p2_first = p2[zidx(D, i, j, 0)], p2_last = p2[zidx(D, i, j, nn - 1)];
  const T t1 = mul_rn(D.c[nn], p2_first), t2 = mul_rn(D.a[nn], p2_last);
  const T bl = add_rn(D.b[nn], lam);
  const T den = add_rn(add_rn(bl, t1), t2);
  const T tol = mul_rn(eps_of<T>(), fmax(fabs(bl), fabs(add_rn(t1, t2))));
  den_c[zden(D, i, j)] = (!D.nopin && fabs(den) <= tol) ? T(0) : den;
}

// ---- sequential substitution, reference operation order ---------------------
template <class T>
__global__ void __launch_bounds__(128) thomas_seq_kernel(const ThomasDev<T> D, const CacheState* st, const T* zbase, const T* p2base,
                                                         const T* denbase, long long slot_z, long long slot_den, T* p, T norm) {
  const int sel = st->sel;
  const T* z = zbase + (long long)sel * slot_z;
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)D.nx * D.ny) return;
  const int j = (int)(col / D.nx), i = (int)(col - (long long)j * D.nx);
  T* pc = p + (long long)j * D.sj + i;
  const int nn = D.nn;
  T pv = mul_rn(mul_rn(pc[0], norm), z[zidx(D, i, j, 0)]);
  pc[0] = pv;
  for (int k = 1; k < nn; ++k) {
    const T zz = z[zidx(D, i, j, k)];
    pv = mul_rn(sub_rn(mul_rn(pc[(long long)k * D.sk], norm), mul_rn(D.a[k], pv)), zz);
    pc[(long long)k * D.sk] = pv;
  }
  for (int k = nn - 2; k >= 0; --k) {
    const T dk = mul_rn(D.c[k], z[zidx(D, i, j, k)]);
    pv = sub_rn(pc[(long long)k * D.sk], mul_rn(dk, pv));
    pc[(long long)k * D.sk] = pv;
  }
  if (!D.periodic) return;
  const T* p2 = p2base + (long long)sel * slot_z;
  const T den = denbase[(long long)sel * slot_den + zden(D, i, j)];
  const T p_first = pc[0], p_last = pc[(long long)(nn - 1) * D.sk];
  T num = sub_rn(sub_rn(mul_rn(pc[(long long)nn * D.sk], norm), mul_rn(D.c[nn], p_first)), mul_rn(D.a[nn], p_last));
  const T pcl = (den == T(0)) ? T(0) : div_rn(num, den);
  pc[(long long)nn * D.sk] = pcl;
  for (int k = 0; k < nn; ++k) pc[(long long)k * D.sk] = add_rn(pc[(long long)k * D.sk], mul_rn(p2[zidx(D, i, j, k)], pcl));
}

// ---- pipelined chunked substitution ---------------------------------------------
// Persistent kernel, one CTA of 1024 threads per SM.  A tile is COLS consecutive
// columns (one 128-byte row segment of the z pencil for COLS = 16 in FP64, 64 bytes
// for COLS = 8) x all rows; thread (c, g) owns the m = ceil(nn / CHUNKS) consecutive
// rows g*m .. g*m+m-1 of column c.  COLS x CHUNKS = 1024 threads always:
//   nz <= 512  : 16 columns x 64 chunks  (m <= 8)
//   nz <= 1024 :  8 columns x 128 chunks (m <= 8), same shared-memory footprint
//   * the tile of pivots and right-hand sides arrives by 16-byte cp.async into
//     shared memory; every thread then keeps its m values and pivots in registers;
//   * while tile i is being solved, tile i+1 is already in flight to the (free
//     again) shared-memory tile, so HBM latency is hidden by the pipeline, not by
//     occupancy.
// Both first-order recurrences are evaluated chunk-parallel: local sweep ->
// (value, product) of every chunk to shared memory -> one warp per column folds
// the CHUNKS chunk maps with a shuffle scan -> fix-up.  Nothing but the final result
// is written: traffic = read p + read z + write p = 24 B/point (FP64).
#define CB_TH_THREADS 1024

__device__ __forceinline__ void cp_async_elem(double* dst_smem, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float* dst_smem, const float* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared memory: pivot tile [CHUNKS m rows][COLS], right-hand-side tile likewise, a/c tables, fold arrays
template <class T, int MMAX, int COLS> constexpr size_t thomas_pipe_smem() {
  return ((size_t)2 * MMAX * CB_TH_THREADS + 2 * (CB_TH_THREADS / COLS) * (MMAX + 2) + 2 * COLS * (CB_TH_THREADS / COLS + 1)) * sizeof(T);
}
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

__device__ __forceinline__ double shfl_up_t(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ float shfl_up_t(float v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_down_t(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ float shfl_down_t(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers: one elected thread fetches a whole tile ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// bounded wait: a byte-count mismatch must trap, not hang the device
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  for (unsigned it = 0;; ++it) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (it > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- thread-block cluster helpers (two CTAs share one column tile when nz > 512) ----
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v into the shared memory of CTA `rank` of the cluster, at the address that `local` has in this CTA
template <class T> __device__ __forceinline__ void dsmem_store(T* local, unsigned rank, T v);
template <> __device__ __forceinline__ void dsmem_store<double>(double* local, unsigned rank, double v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
template <> __device__ __forceinline__ void dsmem_store<float>(float* local, unsigned rank, float v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// one-way hand-over: store v into CTA `rank`'s copy of `local` and complete 8 / 4 bytes of the transaction count of
// ITS copy of the mbarrier `bar` (st.async): the receiver waits on its own barrier, nobody fences global memory
__device__ __forceinline__ void dsmem_send(double* local, unsigned long long* bar, unsigned rank, double v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local), lb = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(ra), "l"(__double_as_longlong(v)), "r"(rb) : "memory");
}
__device__ __forceinline__ void dsmem_send(float* local, unsigned long long* bar, unsigned rank, float v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local), lb = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  for (unsigned it = 0;; ++it) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (it > (1u << 26)) __trap();
  }
}

// Fold of the CHUNKS chunk maps v -> Y_g + P_g v of every column: warp w < COLS owns column w, lane l owns
// the CPL = CHUNKS / 32 consecutive chunks l*CPL ..; Kogge-Stone scan of the lane aggregates with shuffles;
// result = value entering each chunk, written over sY.  FWD: chunk 0 upward; otherwise the last chunk downward.
// `cin` (indexed by column) is the value entering the first chunk in application order (nullptr = 0); the value
// leaving the last one is handed to `agg(column, value)`.
// PEX: also leave in sP the product of the maps before each chunk (the sensitivity of the value entering the chunk
// to a carry entering the CTA), so that a carry that arrives later is applied as vin += pex * carry.
template <class T, bool FWD, int COLS, int CHUNKS, bool PEX, class Agg>
__device__ __forceinline__ void thomas_fold(T* sY, T* sP, int tid, const T* cin, Agg agg) {
  constexpr int LD = CHUNKS + 1, CPL = CHUNKS / 32;
  static_assert(CHUNKS % 32 == 0, "chunks per column must be a multiple of the warp size");
  const int w = tid >> 5, lane = tid & 31;
  if (w < COLS) {
    T* rowY = sY + w * LD + lane * CPL;
    T* rowP = sP + w * LD + lane * CPL;
    T y[CPL], p[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) { y[q] = rowY[q]; p[q] = rowP[q]; }
    // aggregate map of the lane, in application order
    T Y = FWD ? y[0] : y[CPL - 1];
    T P = FWD ? p[0] : p[CPL - 1];
#pragma unroll
    for (int s = 1; s < CPL; ++s) {
      const int q = FWD ? s : CPL - 1 - s;
      Y = fma(p[q], Y, y[q]);
      P *= p[q];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const T Yp = FWD ? shfl_up_t(Y, d) : shfl_down_t(Y, d);
      const T Pp = FWD ? shfl_up_t(P, d) : shfl_down_t(P, d);
      const bool on = FWD ? (lane >= d) : (lane + d < 32);
      if (on) { Y = fma(P, Yp, Y); P *= Pp; }
    }
    const T c0 = cin ? cin[w] : T(0);
    const bool last = FWD ? (lane == 31) : (lane == 0);
    if (last) agg(w, fma(P, c0, Y));
    T vin = FWD ? shfl_up_t(Y, 1) : shfl_down_t(Y, 1);
    T pex = FWD ? shfl_up_t(P, 1) : shfl_down_t(P, 1);
    if (FWD ? (lane == 0) : (lane == 31)) { vin = c0; pex = T(1); }
    else vin = fma(pex, c0, vin);
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
      const int q = FWD ? s : CPL - 1 - s;
      rowY[q] = vin;
      if (PEX) rowP[q] = pex;
      vin = fma(p[q], vin, y[q]);
      pex *= p[q];
    }
  }
}

// EXACT: the rows per chunk equal MMAX (no per-row predicates in the sweeps).
// VEC: columns come in aligned 16-byte groups, so the tile copies use 16-byte cp.async.
// CL = 2: a cluster of two CTAs shares one tile of COLS columns; CTA `rank` owns the rows
// [rank CHUNKS m, (rank + 1) CHUNKS m) and the two folds hand their carry across through distributed shared
// memory (forward: rank 0 -> 1, backward: rank 1 -> 0), so 16-column (128-byte) rows reach nz = 1024.
// LDM: how a tile reaches shared memory.  0 = element-wise cp.async (any alignment), 1 = 16-byte cp.async,
// 2 = TMA: one thread issues cp.async.bulk.tensor boxes of {COLS columns x box_rows rows} of the right-hand side
// (3-D map: x, y, row) and of the cached pivots (4-D map: x, y, row, cache slot), completion on an mbarrier;
// out-of-range rows / columns arrive as zeros (the identity chunk maps the padding needs).
#define CB_TH_LD_ELEM 0
#define CB_TH_LD_VEC 1
#define CB_TH_LD_TMA 2
// TALL: tiles of several y rows (ThomasDev::jb) -- a compile-time switch, so that the one-row-per-tile instantiations keep their
// register budget (the run-time version spilled 24 bytes at the 64-register cap: C3 0.95 -> 1.02 ms).
template <class T, int MMAX, bool EXACT, int LDM, int COLS, int CL, bool TALL = false>
__global__ void __launch_bounds__(CB_TH_THREADS, 1)
thomas_pipe_kernel(const ThomasDev<T> D, const CacheState* st, const T* zbase, const T* p2base, const T* denbase,
                   long long slot_z, long long slot_den, T* p, T norm, const __grid_constant__ CUtensorMap map_p,
                   const __grid_constant__ CUtensorMap map_z, int box_rows) {
  constexpr bool VEC = LDM == CB_TH_LD_VEC;
  constexpr bool TMA = LDM == CB_TH_LD_TMA;
  constexpr int NT = CB_TH_THREADS;
  constexpr int CHUNKS = NT / COLS;
  constexpr int LD = CHUNKS + 1;
  constexpr int CSA = MMAX + 2;                     // chunk stride of the a/c tables (16-byte aligned, bank spread)
  constexpr int VW = 16 / sizeof(T);                // elements per 16-byte piece
  extern __shared__ __align__(128) unsigned char cb_smem_raw[];
  T* zs = reinterpret_cast<T*>(cb_smem_raw);        // [row][COLS] pivots of the tile in flight / being consumed
  T* ps = zs + (size_t)MMAX * NT;                   // [row][COLS] right-hand side
  T* sa = ps + (size_t)MMAX * NT;                   // a_k at [g][r]
  T* sc = sa + CHUNKS * CSA;                        // c_k
  T* sY = sc + CHUNKS * CSA;                        // [c][g] chunk value, row stride CHUNKS + 1
  T* sP = sY + COLS * LD;                           // [c][g] chunk product
  __shared__ T xchg[2][COLS];                       // carries handed over by the other CTA of the cluster
  __shared__ __align__(8) unsigned long long xbar[2];   // their arrival: [0] forward carry (rank 1 waits), [1] backward (rank 0)
  __shared__ __align__(8) unsigned long long tile_bar;   // TMA: completion barrier of the tile in flight
  const unsigned rank = CL > 1 ? cluster_ctarank() : 0u;
  const int tid = threadIdx.x, c = tid & (COLS - 1), g = tid / COLS;
  const int jb = TALL ? D.jb : 1;
  const int nn1 = D.nn;                      // rows of one system
  const int m = EXACT ? MMAX : D.m, nn = nn1 * jb;   // rows of one tile
  const int nrows_tile = CHUNKS * m;
  const long long ncol = (long long)D.nx * D.ny;
  const long long sk = D.sk;
  const int tiles_x = (D.xn + COLS - 1) / COLS;
  const int nyt = D.ny / jb;                 // tiles along y
  const int ntiles = tiles_x * nyt;
  const int sel = st->sel;
  const int rbase = (int)rank * nrows_tile;         // first row of this CTA
  const int k0 = rbase + g * m;
  const int nrow_full = nn - k0 < 0 ? 0 : (nn - k0 < m ? nn - k0 : m);   // my rows inside the system
  const T* zsel = zbase + (long long)sel * slot_z;

  for (int e = tid; e < nrows_tile; e += NT) {
    const int ge = e / m, re = e - ge * m;
    const int kk = TALL ? (rbase + e) % nn1 : rbase + e;   // row inside its own system
    T av = (rbase + e < nn) ? D.a[kk] : T(0);
    T cv = (rbase + e < nn) ? D.c[kk] : T(0);
    if (TALL) {        // seams between the systems of a tall tile
      if (kk == 0) av = T(0);
      if (kk == nn1 - 1) cv = T(0);
    }
    if (D.dt_mode) {   // distributed TDMA: two chain starts, two untouched last rows (see ThomasDev::dt_mode)
      if (rbase + e == 1) av = T(0);
      if (rbase + e >= nn - 2) cv = T(0);
    }
    sa[ge * CSA + re] = av;
    sc[ge * CSA + re] = cv;
  }

  if (TMA) {
    if (tid == 0) {
      mbar_init(&tile_bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
  }
  unsigned tile_phase = 0;

  // tile number -> (row j of the z pencil, x tile).  Deduplicated cache on a full-width launch: the (up to four) tiles that
  // share their pivots get consecutive tile numbers, so that neighbouring CTAs of the persistent grid fetch the same pivot
  // tile at the same time -- one HBM read, the others hit L2.  y pairs in the order (0, ny/2), (1, ny-1), (2, ny-2), ...
  const bool grouped = (D.dx | D.dy) && D.xb == 0 && D.xn == D.nx && (tiles_x & 1) == 0 && (!D.dy || (D.ny & 1) == 0);
  auto decode = [&](int t, int& tj, int& ti) {   // tj = FIRST y row of the tile
    if (!grouped) { tj = t / tiles_x; ti = t - tj * tiles_x; tj *= jb; return; }
    int mx = 0, my = 0;
    if (D.dx) { mx = t & 1; t >>= 1; }
    if (D.dy) { my = t & 1; t >>= 1; }
    const int ntu = D.dx ? tiles_x / 2 : tiles_x;
    const int qy = t / ntu, tu = t - qy * ntu;
    ti = mx ? tu + tiles_x / 2 : tu;
    tj = !D.dy ? qy * jb : (qy == 0 ? (my ? D.ny / 2 : 0) : (my ? D.ny - qy : qy));   // (the host never combines dy with jb > 1)
  };
  // (flat column of the tile's first thread, live columns); the field itself may have a row pitch sj != nx
  // (the haloed array of solver_gaussel_z): its columns sit at pcol = col + tj (sj - nx)
  const long long pitch_extra = D.sj - D.nx;
  auto tile_col0 = [&](int tj, int ti, int& ncols) -> long long {
    const int ti0 = ti * COLS;
    ncols = D.xn - ti0 < COLS ? D.xn - ti0 : COLS;
    return (long long)tj * D.nx + D.xb + ti0;
  };
  // pivots of the tile whose first column is x0: stored row of the cache and first stored column (tiles never straddle nxu:
  // both are multiples of the tile width)
  auto ztile = [&](int tj, int x0, int& zx0) -> int {
    zx0 = zux(D, x0);
    return zuy(D, tj);
  };
  // asynchronous copy of one tile of pivots and right-hand sides into [row][COLS]; rows >= nn and dead
  // columns are zero filled so that the chunk maps of padding rows are exact identities / zeros
  auto prefetch = [&](int tj, int ti, long long col0, int ncols) {
    const int x0 = D.xb + ti * COLS;
    int zx0;
    const int ju = ztile(tj, x0, zx0);
    const long long pshift = (long long)tj * pitch_extra;
    if (TMA) {
      if (tid == 0) {
        fence_proxy_async();   // the generic-proxy reads of the previous tile are ordered before the async writes
        mbar_expect_tx(&tile_bar, (unsigned)(2u * nrows_tile * COLS * sizeof(T)));
        for (int r0 = 0; r0 < nrows_tile; r0 += box_rows) {
          if (TALL) {     // maps of a tall tile describe the (j, k) rows as one run: (x, 0, j nn + k) / (x, j nn + k, 0, slot)
            tma_load_3d(ps + r0 * COLS, &map_p, &tile_bar, x0, 0, tj * nn1 + rbase + r0);
            tma_load_4d(zs + r0 * COLS, &map_z, &tile_bar, zx0, ju * nn1 + rbase + r0, 0, sel);
          } else {
            tma_load_3d(ps + r0 * COLS, &map_p, &tile_bar, x0, tj, rbase + r0);
            tma_load_4d(zs + r0 * COLS, &map_z, &tile_bar, zx0, rbase + r0, ju, sel);
          }
        }
      }
      return;
    }
    const T* zrow0 = zsel + (long long)ju * D.zsj;   // stored row ju, k = 0
    if (VEC) {
      constexpr int PPR = COLS / VW;                // pieces per row
      for (int q = tid; q < nrows_tile * PPR; q += NT) {
        const int row = q / PPR, pc = (q - row * PPR) * VW;
        T* zd = zs + row * COLS + pc;
        T* pd = ps + row * COLS + pc;
        if (rbase + row < nn && pc < ncols) {
          const T* zr = zrow0 + (long long)(rbase + row) * D.zsk;
          cp_async_16(zd, zr + zx0 + pc);
          cp_async_16(pd, p + (long long)(rbase + row) * sk + col0 + pshift + pc);
        } else {
#pragma unroll
          for (int e = 0; e < VW; ++e) { zd[e] = T(0); pd[e] = T(0); }
        }
      }
    } else {
      for (int q = tid; q < nrows_tile * COLS; q += NT) {
        const int row = q / COLS, pc = q - row * COLS;
        if (rbase + row < nn && pc < ncols) {
          cp_async_elem(zs + q, zrow0 + (long long)(rbase + row) * D.zsk + zx0 + pc);
          cp_async_elem(ps + q, p + (long long)(rbase + row) * sk + col0 + pshift + pc);
        } else {
          zs[q] = T(0);
          ps[q] = T(0);
        }
      }
    }
    cp_async_commit();
  };

  const int tstride = (int)gridDim.x / CL;
  int tile = (int)blockIdx.x / CL;
  if (tile >= ntiles) return;   // both CTAs of a cluster leave together
  int ncols, tj, ti;
  decode(tile, tj, ti);
  long long col0 = tile_col0(tj, ti, ncols);
  prefetch(tj, ti, col0, ncols);
  auto no_agg = [](int, T) {};
  unsigned xphase = 0;
  if (CL > 1) {
    // hand-over barriers: armed by the receiver (one arrival + COLS values), completed by the sender's st.async
    if (tid == 0) {
      mbar_init(&xbar[0], 1);
      mbar_init(&xbar[1], 1);
      fence_mbar_init();
      mbar_expect_tx(&xbar[rank == 1 ? 0 : 1], (unsigned)(COLS * sizeof(T)));
    }
    cluster_sync_all();   // both CTAs are armed before either one signals
  }
  for (; tile < ntiles; tile += tstride) {
    if (TMA) {
      mbar_wait(&tile_bar, tile_phase);   // every thread waits itself: the tile is visible to it without a CTA barrier
      tile_phase ^= 1u;
    } else {
      cp_async_wait<0>();
    }
    // cp.async paths: the copies of ALL threads must have landed.  TMA path: the barrier is only needed where threads read
    // fold slots other than their own after the last barrier of the previous tile (periodic closure, CTA pairs).
    if (!TMA || CL > 1 || D.periodic) __syncthreads();
    T y[MMAX], z[MMAX];
#pragma unroll
    for (int r = 0; r < MMAX; ++r)
      if (EXACT || r < m) {
        y[r] = ps[(g * m + r) * COLS + c];
        z[r] = zs[(g * m + r) * COLS + c];
      }
    const bool live = c < ncols;
    const int nrow = live ? nrow_full : 0;
    const long long col = col0 + c;
    const long long pcol = col + (long long)tj * pitch_extra;

    // forward: y_k = (p_k norm - a_k y_{k-1}) z_k, chunk-local with y_{k0-1} := 0
    T yy = T(0), pi = T(1);
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      if (EXACT || r < m) {
        const T al = -(sa[g * CSA + r] * z[r]);
        yy = fma(al, yy, y[r] * norm * z[r]);
        pi *= al;
        y[r] = yy;
      }
    }
    sY[c * LD + g] = yy;
    sP[c * LD + g] = pi;
    __syncthreads();   // every thread holds its tile values in registers: the tile buffers are free again
    const int tnext = tile + tstride;
    int ncols_n = 0, tj_n = 0, ti_n = 0;
    long long col0_n = 0;
    if (tnext < ntiles) {
      decode(tnext, tj_n, ti_n);
      col0_n = tile_col0(tj_n, ti_n, ncols_n);
      prefetch(tj_n, ti_n, col0_n, ncols_n);
    }
    T vin;
    if (CL == 1) {
      thomas_fold<T, true, COLS, CHUNKS, false>(sY, sP, tid, (const T*)nullptr, no_agg);
      __syncthreads();
      vin = sY[c * LD + g];
    } else {
      // both CTAs fold their halves at once with a zero carry; rank 0 hands the value leaving its last row to
      // rank 1 (one-way st.async), which adds (product of the maps before the chunk) x carry
      if (rank == 0) thomas_fold<T, true, COLS, CHUNKS, false>(sY, sP, tid, (const T*)nullptr, [&](int w, T v) { dsmem_send(&xchg[0][w], &xbar[0], 1u, v); });
      else thomas_fold<T, true, COLS, CHUNKS, true>(sY, sP, tid, (const T*)nullptr, no_agg);
      __syncthreads();
      vin = sY[c * LD + g];
      if (rank == 1) {
        mbar_wait_cluster(&xbar[0], xphase);
        vin = fma(sP[c * LD + g], xchg[0][c], vin);
        // re-arm for the next tile: rank 0 cannot send again before it has received this tile's backward carry
        if (tid == 0) mbar_expect_tx(&xbar[0], (unsigned)(COLS * sizeof(T)));
      }
    }
    pi = T(1);
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      if (EXACT || r < m) {
        pi *= -(sa[g * CSA + r] * z[r]);
        y[r] = fma(pi, vin, y[r]);
      }
    }
    // backward: x_k = y_k - d_k x_{k+1}, d_k = c_k z_k
    T xx = T(0), rho = T(1);
#pragma unroll
    for (int r = MMAX - 1; r >= 0; --r) {
      if (EXACT || r < m) {
        const T nd = -(sc[g * CSA + r] * z[r]);
        xx = fma(nd, xx, y[r]);
        rho *= nd;
        y[r] = xx;
      }
    }
    sY[c * LD + g] = xx;   // own slot: only this thread read it since the fold
    sP[c * LD + g] = rho;
    __syncthreads();
    if (CL == 1) {
      thomas_fold<T, false, COLS, CHUNKS, false>(sY, sP, tid, (const T*)nullptr, no_agg);
      __syncthreads();
      vin = sY[c * LD + g];
    } else {
      // the same downward: rank 1 hands the value leaving its first row to rank 0
      if (rank == 1) thomas_fold<T, false, COLS, CHUNKS, false>(sY, sP, tid, (const T*)nullptr, [&](int w, T v) { dsmem_send(&xchg[1][w], &xbar[1], 0u, v); });
      else thomas_fold<T, false, COLS, CHUNKS, true>(sY, sP, tid, (const T*)nullptr, no_agg);
      __syncthreads();
      vin = sY[c * LD + g];
      if (rank == 0) {
        mbar_wait_cluster(&xbar[1], xphase);
        vin = fma(sP[c * LD + g], xchg[1][c], vin);
        if (tid == 0) mbar_expect_tx(&xbar[1], (unsigned)(COLS * sizeof(T)));
      }
      xphase ^= 1u;
    }
    rho = T(1);
#pragma unroll
    for (int r = MMAX - 1; r >= 0; --r) {
      if (EXACT || r < m) {
        rho *= -(sc[g * CSA + r] * z[r]);
        y[r] = fma(rho, vin, y[r]);
      }
    }
    if (CL == 1 && D.periodic) {
      // closure value and rank-one correction (clusters: the host never pairs CTAs on a periodic system), src/solver.f90:272-306
      __syncthreads();
      const int kl = nn - 1;
      if (g == 0) sY[c] = y[0];                  // x_1
      if (kl >= k0 && kl < k0 + m) {
#pragma unroll
        for (int r = 0; r < MMAX; ++r)
          if (k0 + r == kl) sY[COLS + c] = y[r];   // x_nn
      }
      __syncthreads();
      if (g == 0) {
        T pcl = T(0);
        if (live) {
          const T den = denbase[(long long)sel * slot_den + zden(D, (int)(col - (long long)tj * D.nx), tj)];
          const T pnn = p[(long long)nn * sk + pcol];
          const T num = sub_rn(sub_rn(mul_rn(pnn, norm), mul_rn(D.c[nn], sY[c])), mul_rn(D.a[nn], sY[COLS + c]));
          pcl = (den == T(0)) ? T(0) : div_rn(num, den);
          if (D.out_rows) D.out_rows[nn].ptr[(long long)tj * D.out_rows[nn].sj + (col - (long long)tj * D.nx)] = pcl;
          else p[(long long)nn * sk + pcol] = pcl;
        }
        sP[c] = pcl;
      }
      __syncthreads();
      const T pcl = sP[c];
      const T* p2c = p2base + (long long)sel * slot_z + zidx(D, live ? (int)(col - (long long)tj * D.nx) : 0, tj, k0);
#pragma unroll
      for (int r = 0; r < MMAX; ++r)
        if (r < nrow) y[r] = fma(p2c[(long long)r * D.nxu], pcl, y[r]);
    }
    if (D.dt_mode && live) {
      // first row of the slab: (y_0 - cc_0 x_1) * Z1; first and last row are this rank's rows of the reduced system
      const long long ncolg = (long long)D.nx * D.ny;
      if (k0 == 0) {
        y[0] = y[0] * D.dt_z1[(long long)sel * D.dt_slot_small + col];
        D.dt_rp[col] = y[0];
      }
      if (nn - 1 >= k0 && nn - 1 < k0 + m) {
#pragma unroll
        for (int r = 0; r < MMAX; ++r)
          if (k0 + r == nn - 1) D.dt_rp[ncolg + col] = y[r];
      }
    }
    if (D.out_rows) {
      const OutRow<T>* rows = D.out_rows + k0;
      const long long xi = col - (long long)tj * D.nx;
#pragma unroll
      for (int r = 0; r < MMAX; ++r)
        if (r < nrow) {
          const longlong2 e = __ldg(reinterpret_cast<const longlong2*>(rows + r));
          reinterpret_cast<T*>(e.x)[(long long)tj * e.y + xi] = y[r];
        }
    } else {
      T* pc = p + (long long)k0 * sk + pcol;
      if (nrow == m) {
#pragma unroll
        for (int r = 0; r < MMAX; ++r)
          if (EXACT || r < m) pc[r * sk] = y[r];
      } else {
#pragma unroll
        for (int r = 0; r < MMAX; ++r)
          if (r < nrow) pc[r * sk] = y[r];
      }
    }
    col0 = col0_n;
    ncols = ncols_n;
    tj = tj_n;
    ti = ti_n;
  }
  if (CL > 1) cluster_sync_all();   // leave together: a CTA's shared memory must outlive the peer's last hand-over
}

}  // namespace cb
