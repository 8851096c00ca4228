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
//   thomas_warp_kernel   : one WARP per column; each lane keeps a chunk of m
//       rows (values and pivots) in registers and the two first-order linear
//       recurrences are evaluated chunk-parallel (local sweep, warp-shuffle
//       scan of the affine maps, fix-up).  p' never leaves the chip: traffic
//       is read p + read z + write p = 24 B/pt instead of 48.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

template <class T> struct ThomasDev {
  int nx, ny;         // columns (x fastest)
  int n;              // rows handled: nz - q, incl. the periodic closure row
  int nn;             // n-1 if periodic else n
  int periodic;
  long long sj, sk;   // field strides: p[k*sk + j*sj + i]
  const T* a; const T* b; const T* c;  // device, length >= n
  const T* lam; long long lam_sj;      // lambdaxy[j*lam_sj + i]
  int m;              // rows per lane in the chunk layout = ceil(nn/32)
  int chunk_layout;   // 0: z[k][j][i]   1: z[((j*nx+i)*m + r)*32 + lane], k = lane*m + r
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

template <class T> __device__ __forceinline__ long long zidx(const ThomasDev<T>& D, int i, int j, int k) {
  if (D.chunk_layout) {
    const int l = k / D.m, r = k - l * D.m;
    return (((long long)j * D.nx + i) * D.m + r) * 32 + l;
  }
  return ((long long)k * D.ny + j) * D.nx + i;
}

// ---- factorisation cache bookkeeping (device resident) ----------------------
#define CB_MAX_SLOTS 8
struct CacheState {
  unsigned long long key_new;
  unsigned long long keys[CB_MAX_SLOTS];
  unsigned long long stamp[CB_MAX_SLOTS];
  unsigned long long clock;
  unsigned long long nfactor;  // how many factorisations ran (diagnostic)
  int sel, hit, nslots, pad;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long bits_of(float v) { return (unsigned long long)__float_as_uint(v); }

// order-independent content hash of (a, b, c, lambdaxy)
template <class T>
__global__ void thomas_hash_kernel(const ThomasDev<T> D, CacheState* st) {
  const long long ncoef = D.n;
  const long long nlam = (long long)D.nx * D.ny;
  const long long total = 3 * ncoef + nlam;
  unsigned long long h = 0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    T v;
    if (e < ncoef) v = D.a[e];
    else if (e < 2 * ncoef) v = D.b[e - ncoef];
    else if (e < 3 * ncoef) v = D.c[e - 2 * ncoef];
    else {
      const long long q = e - 3 * ncoef;
      const long long j = q / D.nx, i = q - j * D.nx;
      v = D.lam[j * D.lam_sj + i];
    }
    h += mix64(mix64((unsigned long long)e) ^ bits_of(v));
  }
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&st->key_new, h);
}

__global__ void thomas_select_kernel(CacheState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long key = st->key_new | 1ULL;  // never 0 (0 = empty slot)
  st->clock += 1;
  int sel = -1;
  for (int s = 0; s < st->nslots; ++s)
    if (st->keys[s] == key) sel = s;
  if (sel >= 0) {
    st->hit = 1;
  } else {
    st->hit = 0;
    sel = 0;
    for (int s = 1; s < st->nslots; ++s)
      if (st->stamp[s] < st->stamp[sel]) sel = s;
    st->keys[sel] = key;
    st->nfactor += 1;
  }
  st->stamp[sel] = st->clock;
  st->sel = sel;
  st->key_new = 0;
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
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)D.nx * D.ny) return;
  const int j = (int)(col / D.nx), i = (int)(col - (long long)j * D.nx);
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
    if (k == nn - 1) {
      const T tol = mul_rn(eps_of<T>(), fmax(fabs(bl), fabs(ad)));
      pin = fabs(den) <= tol;
    }
    zz = pin ? T(0) : div_rn(one, den);
    d = mul_rn(D.c[k], zz);
    z[zidx(D, i, j, k)] = zz;
  }
  // pad the chunk layout so that lanes past nn see identity rows
  if (D.chunk_layout)
    for (int k = nn; k < 32 * D.m; ++k) z[zidx(D, i, j, k)] = T(0);
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
  if (D.chunk_layout)
    for (int k = nn; k < 32 * D.m; ++k) p2[zidx(D, i, j, k)] = T(0);
  const T p2_first = p2[zidx(D, i, j, 0)], p2_last = p2[zidx(D, i, j, nn - 1)];
  const T t1 = mul_rn(D.c[nn], p2_first), t2 = mul_rn(D.a[nn], p2_last);
  const T bl = add_rn(D.b[nn], lam);
  const T den = add_rn(add_rn(bl, t1), t2);
  const T tol = mul_rn(eps_of<T>(), fmax(fabs(bl), fabs(add_rn(t1, t2))));
  den_c[col] = (fabs(den) <= tol) ? T(0) : den;
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
  const T den = denbase[(long long)sel * slot_den + col];
  const T p_first = pc[0], p_last = pc[(long long)(nn - 1) * D.sk];
  T num = sub_rn(sub_rn(mul_rn(pc[(long long)nn * D.sk], norm), mul_rn(D.c[nn], p_first)), mul_rn(D.a[nn], p_last));
  const T pcl = (den == T(0)) ? T(0) : div_rn(num, den);
  pc[(long long)nn * D.sk] = pcl;
  for (int k = 0; k < nn; ++k) pc[(long long)k * D.sk] = add_rn(pc[(long long)k * D.sk], mul_rn(p2[zidx(D, i, j, k)], pcl));
}

// ---- warp-per-column chunked substitution ------------------------------------
// shared tile: column c occupies KP elements; row k = l*m + r sits at l*CS + r,
// CS = m | 1 (odd -> the 32 lanes of a column hit distinct banks), KP = 2 mod 16
// (-> the cooperative 8-wide row copies are conflict free as well).
__host__ __device__ inline int thomas_cs(int m) { return m | 1; }
__host__ __device__ inline int thomas_kp(int m) {
  int kp = 32 * thomas_cs(m) + 1;  // +1: room for the periodic closure row (k = nn <= 32 m)
  while ((kp & 15) != 2) ++kp;
  return kp;
}

template <class T, int MMAX, int CX>
__global__ void __launch_bounds__(32 * CX, (MMAX <= 16 ? 2 : 1))
thomas_warp_kernel(const ThomasDev<T> D, const CacheState* st, const T* zbase, const T* p2base, const T* denbase,
                   long long slot_z, long long slot_den, T* p, T norm) {
  extern __shared__ __align__(16) unsigned char cb_smem_raw[];
  T* s = reinterpret_cast<T*>(cb_smem_raw);
  const int m = D.m, nn = D.nn, n = D.n;
  const int CS = thomas_cs(m), KP = thomas_kp(m);
  T* sa = s + CX * KP;       // -a_k in [r][lane] order
  T* sc = sa + 32 * MMAX;    //  c_k in [r][lane] order
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int tiles_x = (D.nx + CX - 1) / CX;
  const int j = blockIdx.x / tiles_x;
  const int i0 = (blockIdx.x - j * tiles_x) * CX;
  const int sel = st->sel;
  const bool live = (i0 + w) < D.nx;
  const long long col = (long long)j * D.nx + i0 + w;

  // pivots of my chunk straight into registers (coalesced: lane fastest)
  T z[MMAX];
  {
    const T* zc = zbase + (long long)sel * slot_z + col * m * 32 + lane;
#pragma unroll
    for (int r = 0; r < MMAX; ++r) z[r] = (live && r < m) ? zc[(long long)r * 32] : T(0);
  }
  for (int e = tid; e < 32 * m; e += 32 * CX) {
    const int r = e >> 5, l = e & 31, k = l * m + r;
    sa[e] = (k < nn) ? -D.a[k] : T(0);
    sc[e] = (k < nn) ? D.c[k] : T(0);
  }
  // cooperative tile load, rows 0..n-1
  T* pg = p + (long long)j * D.sj + i0;
  for (int e = tid; e < n * CX; e += 32 * CX) {
    const int k = e / CX, c = e - k * CX;
    const int l = k / m, r = k - l * m;
    if (i0 + c < D.nx) s[c * KP + l * CS + r] = pg[(long long)k * D.sk + c];
  }
  __syncthreads();

  T y[MMAX];
  if (live) {
    T* sp = s + w * KP + lane * CS;
    // pass A: local forward sweep  y_k = beta_k + alpha_k y_{k-1}
    T yy = T(0), pi = T(1);
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      if (r < m) {
        const T zz = z[r];
        const T al = sa[r * 32 + lane] * zz;
        const T be = sp[r] * norm * zz;
        yy = fma(al, yy, be);
        pi *= al;
      }
      y[r] = yy;
    }
    // inclusive scan of the affine maps v -> Y + Pi v over the lanes
    T Y = yy, Pi = pi;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const T Yp = __shfl_up_sync(0xffffffffu, Y, dlt);
      const T Pp = __shfl_up_sync(0xffffffffu, Pi, dlt);
      if (lane >= dlt) { Y = fma(Pi, Yp, Y); Pi *= Pp; }
    }
    T Yin = __shfl_up_sync(0xffffffffu, Y, 1);
    if (lane == 0) Yin = T(0);
    // pass B: fix-up with the incoming value
    pi = T(1);
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      if (r < m) {
        pi *= sa[r * 32 + lane] * z[r];
        y[r] = fma(pi, Yin, y[r]);
      }
    }
    // pass C: local backward sweep  x_k = y_k - d_k x_{k+1}
    T xx = T(0), rho = T(1);
#pragma unroll
    for (int r = MMAX - 1; r >= 0; --r) {
      if (r < m) {
        const T nd = -(sc[r * 32 + lane] * z[r]);
        xx = fma(nd, xx, y[r]);
        rho *= nd;
        y[r] = xx;
      }
    }
    T X = xx, R = rho;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const T Xp = __shfl_down_sync(0xffffffffu, X, dlt);
      const T Rp = __shfl_down_sync(0xffffffffu, R, dlt);
      if (lane + dlt < 32) { X = fma(R, Xp, X); R *= Rp; }
    }
    T Xin = __shfl_down_sync(0xffffffffu, X, 1);
    if (lane == 31) Xin = T(0);
    // pass D: fix-up
    rho = T(1);
#pragma unroll
    for (int r = MMAX - 1; r >= 0; --r) {
      if (r < m) {
        rho *= -(sc[r * 32 + lane] * z[r]);
        y[r] = fma(rho, Xin, y[r]);
      }
    }
    if (D.periodic) {
      // closure value and Sherman-Morrison correction, src/solver.f90:272-306
      const int ls = (nn - 1) / m, rs = (nn - 1) - ls * m;
      T v = T(0);
#pragma unroll
      for (int r = 0; r < MMAX; ++r) if (r == rs) v = y[r];
      const T x_last = __shfl_sync(0xffffffffu, v, ls);
      const T x_first = __shfl_sync(0xffffffffu, y[0], 0);
      const T den = denbase[(long long)sel * slot_den + col];
      const int lq = nn / m, rq = nn - lq * m;
      const T pnn = s[w * KP + lq * CS + rq];
      const T num = sub_rn(sub_rn(mul_rn(pnn, norm), mul_rn(D.c[nn], x_first)), mul_rn(D.a[nn], x_last));
      const T pcl = (den == T(0)) ? T(0) : div_rn(num, den);
      const T* p2c = p2base + (long long)sel * slot_z + col * m * 32 + lane;
#pragma unroll
      for (int r = 0; r < MMAX; ++r)
        if (r < m) y[r] = fma(p2c[(long long)r * 32], pcl, y[r]);
      __syncwarp();
      if (lane == 0) s[w * KP + lq * CS + rq] = pcl;
      __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < MMAX; ++r)
      if (r < m && lane * m + r < nn) sp[r] = y[r];
  }
  __syncthreads();
  for (int e = tid; e < n * CX; e += 32 * CX) {
    const int k = e / CX, c = e - k * CX;
    const int l = k / m, r = k - l * m;
    if (i0 + c < D.nx) pg[(long long)k * D.sk + c] = s[c * KP + l * CS + r];
  }
}

}  // namespace cb
