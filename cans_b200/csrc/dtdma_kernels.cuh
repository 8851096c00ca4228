// Distributed TDMA -- the arithmetic of `gaussel_dtdma` / `gaussel_dtdma_gpu`
// (/root/reference/src/solver.f90:309-517, src/solver_gpu.f90:430-695; `is_poisson_dtdma`): the z direction
// stays decomposed, every rank eliminates its inner rows so that they only couple to its first and last row,
// the 2 P boundary rows form a reduced tridiagonal system, and the inner rows are updated from its solution.
//
// The four kernels below are the complete arithmetic, in the reference's operation order (exactly rounded, never
// contracted; the tests compare them with the CPU restatement of the same routine).  Two callers (capi.cu):
// cansb200_gaussel_dtdma keeps the P slabs on ONE GPU ("virtual ranks", stage-level parity), and solve_dist_dtdma
// (CANSB200_CTX_DTDMA) runs the elimination per rank and gathers the two boundary rows of every rank on all peers
// with dtdma_gather_kernel (peer-mapped stores) before the redundant reduced solve.
// One thread per (i, j) column, i fastest: every access is coalesced.
#pragma once
#include <cuda_runtime.h>
#include "thomas_kernels.cuh"   // exactly rounded arithmetic helpers

namespace cb {

#define CB_DTDMA_MAX_RANKS 16

template <class T> struct DtdmaDev {
  int nx, ny;            // columns
  int n;                 // rows of the global system (nz - q)
  int nranks;
  int starts[CB_DTDMA_MAX_RANKS + 1];   // rank r owns rows starts[r] .. min(starts[r+1], n) - 1
  int periodic;
  const T* a; const T* b; const T* c;   // global coefficient arrays (every rank uses its slice)
  const T* lam;                         // lambdaxy[j * nx + i] or nullptr
  T* Z;  T* AA; T* CC;                  // [k][j][i]: elimination pivots, and the couplings of row k to the rank's first / last row
  T* Z1;                                // [rank][j][i]: 1 / (1 - aa_2 cc_1) of the rank's first row
  T* ra; T* rc; T* rcw; T* rp; T* rp2;  // reduced system [2 P][j][i]: coefficients, work copy, right-hand side, periodic auxiliary
  // Coefficient cache of the distributed solve (the reference keeps aa_z / cc_z between calls while the coefficients do not
  // change: is_dtdma_update, src/solver_gpu.f90:571-591).  Z / AA / CC / Z1 / ra / rc are per cache slot; `st` is the plan's
  // factorisation-cache state (content hash of a, b, c, lambda -> slot, hit / miss, all on the device): on a hit the
  // coefficient kernel and the coefficient rows of the gather do nothing.  st = nullptr: no cache (stage-level entry point).
  const CacheState* st;
  long long slot_big, slot_small;       // elements between the slots of (Z, AA, CC) and of (Z1, ra, rc)
};
template <class T> __device__ __forceinline__ long long dtdma_sel_big(const DtdmaDev<T>& D) { return D.st ? (long long)D.st->sel * D.slot_big : 0; }
template <class T> __device__ __forceinline__ long long dtdma_sel_small(const DtdmaDev<T>& D) { return D.st ? (long long)D.st->sel * D.slot_small : 0; }

template <class T> __device__ __forceinline__ void dtdma_range(const DtdmaDev<T>& D, int r, int& k0, int& nl) {
  k0 = D.starts[r];
  const int k1 = D.starts[r + 1] < D.n ? D.starts[r + 1] : D.n;
  nl = k1 - k0;
}

// ---- coefficients: aa, cc of every row and the reduced coefficient rows (src/solver.f90:351-391 without p) ----
template <class T>
__global__ void __launch_bounds__(128) dtdma_coef_kernel(const DtdmaDev<T> D) {
  if (D.st && D.st->hit) return;   // the slot already holds the coefficients of this (a, b, c, lambda)
  const long long ncol = (long long)D.nx * D.ny;
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const T lam = D.lam ? D.lam[col] : T(0);
  const T one = T(1);
  const long long ob = dtdma_sel_big(D), os = dtdma_sel_small(D);
  for (int r = 0; r < D.nranks; ++r) {
    int k0, nl;
    dtdma_range(D, r, k0, nl);
    T* Z = D.Z + ob + (long long)k0 * ncol + col;
    T* AA = D.AA + ob + (long long)k0 * ncol + col;
    T* CC = D.CC + ob + (long long)k0 * ncol + col;
    const T *a = D.a + k0, *b = D.b + k0, *c = D.c + k0;
    for (int k = 0; k < 2; ++k) {
      const T zz = div_rn(one, add_rn(b[k], lam));
      Z[k * ncol] = zz;
      AA[k * ncol] = mul_rn(a[k], zz);
      CC[k * ncol] = mul_rn(c[k], zz);
    }
    T aap = AA[ncol], ccp = CC[ncol];
    for (int k = 2; k < nl; ++k) {   // elimination of lower diagonals
      const T z = div_rn(one, sub_rn(add_rn(b[k], lam), mul_rn(a[k], ccp)));
      aap = mul_rn(mul_rn(-a[k], aap), z);
      ccp = mul_rn(c[k], z);
      Z[k * ncol] = z;
      AA[k * ncol] = aap;
      CC[k * ncol] = ccp;
    }
    T aan = AA[(long long)(nl - 2) * ncol], ccn = CC[(long long)(nl - 2) * ncol];
    for (int k = nl - 3; k >= 1; --k) {   // elimination of upper diagonals
      const T cck = CC[k * ncol];
      aan = sub_rn(AA[k * ncol], mul_rn(cck, aan));
      ccn = -mul_rn(cck, ccn);
      AA[k * ncol] = aan;
      CC[k * ncol] = ccn;
    }
    // first row: aan, ccn are now aa(2), cc(2) of the Fortran (row index 1 here)
    const T aa1 = AA[ncol], cc1 = CC[ncol], cc0 = CC[0];
    const T z1 = div_rn(one, sub_rn(one, mul_rn(aa1, cc0)));
    AA[0] = mul_rn(AA[0], z1);
    CC[0] = mul_rn(-mul_rn(cc0, cc1), z1);
    D.Z1[os + (long long)r * ncol + col] = z1;
    D.ra[os + (long long)(2 * r) * ncol + col] = AA[0];
    D.ra[os + (long long)(2 * r + 1) * ncol + col] = AA[(long long)(nl - 1) * ncol];
    D.rc[os + (long long)(2 * r) * ncol + col] = CC[0];
    D.rc[os + (long long)(2 * r + 1) * ncol + col] = CC[(long long)(nl - 1) * ncol];
  }
}

// ---- phase 1: the right-hand side through the same eliminations; boundary rows to the reduced system ----------
template <class T>
__global__ void __launch_bounds__(128) dtdma_phase1_kernel(const DtdmaDev<T> D, T* p, T norm) {
  const long long ncol = (long long)D.nx * D.ny;
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const long long ob = dtdma_sel_big(D), os = dtdma_sel_small(D);
  for (int r = 0; r < D.nranks; ++r) {
    int k0, nl;
    dtdma_range(D, r, k0, nl);
    T* pc = p + (long long)k0 * ncol + col;
    const T* Z = D.Z + ob + (long long)k0 * ncol + col;
    const T *a = D.a + k0, *c = D.c + k0;
    T pv = mul_rn(mul_rn(pc[0], norm), Z[0]);
    pc[0] = pv;
    pv = mul_rn(mul_rn(pc[ncol], norm), Z[ncol]);
    pc[ncol] = pv;
    for (int k = 2; k < nl; ++k) {
      pv = mul_rn(sub_rn(mul_rn(pc[k * ncol], norm), mul_rn(a[k], pv)), Z[k * ncol]);
      pc[k * ncol] = pv;
    }
    pv = pc[(long long)(nl - 2) * ncol];
    for (int k = nl - 3; k >= 1; --k) {
      const T cck = mul_rn(c[k], Z[k * ncol]);   // the row's cc before the upper elimination rewrote it
      pv = sub_rn(pc[k * ncol], mul_rn(cck, pv));
      pc[k * ncol] = pv;
    }
    // pv = p(2) of the Fortran
    const T cc0 = mul_rn(c[0], Z[0]);
    const T p0 = mul_rn(sub_rn(pc[0], mul_rn(cc0, pc[ncol])), D.Z1[os + (long long)r * ncol + col]);
    pc[0] = p0;
    D.rp[(long long)(2 * r) * ncol + col] = p0;
    D.rp[(long long)(2 * r + 1) * ncol + col] = pc[(long long)(nl - 1) * ncol];
  }
}

// ---- reduced system aa_z x_{k-1} + x_k + cc_z x_{k+1} = pp_z (src/solver.f90:449-490) ---------------------------
template <class T>
__global__ void __launch_bounds__(128) dtdma_reduced_kernel(const DtdmaDev<T> D) {
  const long long ncol = (long long)D.nx * D.ny;
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  const T one = T(1);
  const int nr = 2 * D.nranks;
  const int nn = D.periodic ? nr - 1 : nr;
  const long long os = dtdma_sel_small(D);
  const T* ra = D.ra + os + col;
  const T* rc = D.rc + os + col;
  T* cw = D.rcw + col;
  T* rp = D.rp + col;
  cw[0] = rc[0];
  T pv = rp[0], cprev = rc[0];
  for (int k = 1; k < nn; ++k) {
    const T ak = ra[k * ncol];
    const T z = div_rn(one, sub_rn(one, mul_rn(ak, cprev)));
    pv = mul_rn(sub_rn(rp[k * ncol], mul_rn(ak, pv)), z);
    cprev = mul_rn(rc[k * ncol], z);
    rp[k * ncol] = pv;
    cw[k * ncol] = cprev;
  }
  for (int k = nn - 2; k >= 0; --k) {
    pv = sub_rn(rp[k * ncol], mul_rn(cw[k * ncol], pv));
    rp[k * ncol] = pv;
  }
  if (!D.periodic) return;
  // periodic closure: auxiliary system with the ORIGINAL cc_z (cc_z_0), :462-490
  T* p2 = D.rp2 + col;
  T q = -ra[0];
  if (nn == 1) q = sub_rn(q, rc[0]);
  p2[0] = q;
  cprev = rc[0];
  cw[0] = cprev;
  for (int k = 1; k < nn; ++k) {
    const T ak = ra[k * ncol];
    const T z = div_rn(one, sub_rn(one, mul_rn(ak, cprev)));
    const T rhs = (k == nn - 1) ? sub_rn(T(0), rc[k * ncol]) : T(0);
    q = mul_rn(sub_rn(rhs, mul_rn(ak, q)), z);
    cprev = mul_rn(rc[k * ncol], z);
    p2[k * ncol] = q;
    cw[k * ncol] = cprev;
  }
  for (int k = nn - 2; k >= 0; --k) {
    q = sub_rn(p2[k * ncol], mul_rn(cw[k * ncol], q));
    p2[k * ncol] = q;
  }
  const T cl = rc[(long long)nn * ncol], al = ra[(long long)nn * ncol];
  const T num = sub_rn(sub_rn(rp[(long long)nn * ncol], mul_rn(cl, rp[0])), mul_rn(al, rp[(long long)(nn - 1) * ncol]));
  const T den = add_rn(add_rn(one, mul_rn(cl, p2[0])), mul_rn(al, p2[(long long)(nn - 1) * ncol]));
  const T xl = div_rn(num, den);
  rp[(long long)nn * ncol] = xl;
  for (int k = 0; k < nn; ++k) rp[k * ncol] = add_rn(rp[k * ncol], mul_rn(p2[k * ncol], xl));
}

// ---- phase 3: boundary values back, inner rows p_k -= aa_k p_first + cc_k p_last (:499-510) ---------------------
template <class T>
__global__ void dtdma_phase3_kernel(const DtdmaDev<T> D, T* p) {
  const long long ncol = (long long)D.nx * D.ny;
  const long long tot = ncol * D.n;
  const T* AAs = D.AA + dtdma_sel_big(D);
  const T* CCs = D.CC + dtdma_sel_big(D);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / ncol);
    const long long col = e - (long long)k * ncol;
    int r = 0;
    while (r + 1 < D.nranks && k >= D.starts[r + 1]) ++r;
    int k0, nl;
    dtdma_range(D, r, k0, nl);
    const T pf = D.rp[(long long)(2 * r) * ncol + col], pl = D.rp[(long long)(2 * r + 1) * ncol + col];
    const int kl = k - k0;
    if (kl == 0) p[e] = pf;
    else if (kl == nl - 1) p[e] = pl;
    else p[e] = sub_rn(sub_rn(p[e], mul_rn(AAs[e], pf)), mul_rn(CCs[e], pl));
  }
}

// ---- distributed: my two boundary rows of (aa, cc, p) into the gathered reduced system of EVERY rank ---------------
// dst[s] = base of rank s's gather buffer [3][2 P][ncol] (peer mapped); my rows are 2 rank, 2 rank + 1
struct DtdmaPeers { void* dst[CB_DTDMA_MAX_RANKS]; };
// On a cache hit only the right-hand-side rows travel (the receivers kept the coefficient rows of that slot).
template <class T>
__global__ void dtdma_gather_kernel(DtdmaPeers peers, int rank, int nranks, long long ncol, const T* __restrict__ ra,
                                    const T* __restrict__ rc, const T* __restrict__ rp, const CacheState* st, long long slot_small) {
  const long long per = 2 * ncol;                 // my two rows of one array
  const bool hit = st && st->hit;
  const long long os = st ? (long long)st->sel * slot_small : 0;
  const int narr = hit ? 1 : 3;
  const long long tot = narr * per * nranks;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / (narr * per));
    const long long w = e - (long long)s * narr * per;
    const int arr = hit ? 2 : (int)(w / per);
    const long long o = hit ? w : w - (long long)arr * per;   // h * ncol + col
    const T v = arr == 0 ? ra[os + o] : (arr == 1 ? rc[os + o] : rp[o]);
    T* g = reinterpret_cast<T*>(peers.dst[s]);
    g[((long long)arr * 2 * nranks + 2 * rank) * ncol + o] = v;
  }
}

// after the gather of a cache miss: keep the coefficient rows of all ranks in the slot (ga, gc: [2 P][ncol] per slot)
template <class T>
__global__ void dtdma_save_rows_kernel(const T* __restrict__ G, long long n2p_ncol, T* __restrict__ ga, T* __restrict__ gc,
                                       const CacheState* st, long long slot_rows) {
  if (st->hit) return;
  const long long os = (long long)st->sel * slot_rows;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n2p_ncol; e += (long long)gridDim.x * blockDim.x) {
    ga[os + e] = G[e];
    gc[os + e] = G[n2p_ncol + e];
  }
}

}  // namespace cb
