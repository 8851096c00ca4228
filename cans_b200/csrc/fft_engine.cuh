// Batched 1-D real-to-real transform engine (shared-memory, in-place DIF).
//
// Replaces the FFTW3 guru-r2r plans of the reference CPU path
// (/root/reference/src/fft.f90:83-97,148-162,253-255) and the cuFFT R2C/C2R +
// OpenACC Makhoul passes of the reference GPU path (src/fft.f90:369-767) with
// ONE pass per transform: load (+input permutation / pre-twiddle) -> in-place
// mixed-radix complex FFT of length n/2 on packed pairs -> post-twiddle ->
// store, all on a tile of lines staged in shared memory.
//
// Everything here is `__host__ __device__` so that the index arithmetic of
// every phase can be executed thread-by-thread on the CPU (tests/emu_fft.cpp)
// with exactly the code the kernels run.
//
// Output formats are FFTW's (FFTW manual 4.8): halfcomplex for R2HC/HC2R
// (r0..r_{n/2}, i_{ceil(n/2)-1}..i_1), REDFT10/01 = DCT-II/III, RODFT10/01 =
// DST-II/III, all unnormalised.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb {

// FFTW r2r kind numbering (src/fftw.f90:66-85)
enum Kind : int {
  K_R2HC = 0, K_HC2R = 1,
  K_REDFT00 = 3, K_REDFT01 = 4, K_REDFT10 = 5, K_REDFT11 = 6,
  K_RODFT00 = 7, K_RODFT01 = 8, K_RODFT10 = 9, K_RODFT11 = 10
};

CB_HD bool kind_is_forward(int k) { return k == K_R2HC || k == K_REDFT10 || k == K_RODFT10; }
CB_HD bool kind_is_fast(int k) {
  return k == K_R2HC || k == K_HC2R || k == K_REDFT10 || k == K_REDFT01 || k == K_RODFT10 || k == K_RODFT01;
}

template <class T> struct C2 { T x, y; };

template <class T> CB_HD C2<T> cadd(C2<T> a, C2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <class T> CB_HD C2<T> csub(C2<T> a, C2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <class T> CB_HD C2<T> cmul(C2<T> a, C2<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
// multiply by -i  /  +i
template <class T> CB_HD C2<T> mul_mi(C2<T> a) { return {a.y, -a.x}; }
template <class T> CB_HD C2<T> mul_pi(C2<T> a) { return {-a.y, a.x}; }

// ---------------------------------------------------------------------------
// register butterflies, forward sign exp(-2 pi i / R), natural in / natural out
// ---------------------------------------------------------------------------
template <class T> CB_HD void dft2(C2<T>* a) {
  C2<T> t = a[0];
  a[0] = cadd(t, a[1]);
  a[1] = csub(t, a[1]);
}

template <class T> CB_HD void dft3(C2<T>* a) {
  const T c = T(0.86602540378443864676372317075294);
  C2<T> s = cadd(a[1], a[2]), d = csub(a[1], a[2]);
  C2<T> m = {a[0].x - T(0.5) * s.x, a[0].y - T(0.5) * s.y};
  a[0] = cadd(a[0], s);
  a[1] = {m.x + c * d.y, m.y - c * d.x};
  a[2] = {m.x - c * d.y, m.y + c * d.x};
}

template <class T> CB_HD void dft4(C2<T>& a0, C2<T>& a1, C2<T>& a2, C2<T>& a3) {
  C2<T> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = {t1.x + t3.y, t1.y - t3.x};
  a3 = {t1.x - t3.y, t1.y + t3.x};
}
template <class T> CB_HD void dft4(C2<T>* a) { dft4(a[0], a[1], a[2], a[3]); }

template <class T> CB_HD void dft5(C2<T>* a) {
  const T c1 = T(0.30901699437494742410229341718282), c2 = T(-0.80901699437494742410229341718282);
  const T s1 = T(0.95105651629515357211643933337938), s2 = T(0.58778525229247312916870595463907);
  C2<T> p1 = cadd(a[1], a[4]), m1 = csub(a[1], a[4]);
  C2<T> p2 = cadd(a[2], a[3]), m2 = csub(a[2], a[3]);
  C2<T> a0 = a[0];
  a[0] = {a0.x + p1.x + p2.x, a0.y + p1.y + p2.y};
  C2<T> u1 = {a0.x + c1 * p1.x + c2 * p2.x, a0.y + c1 * p1.y + c2 * p2.y};
  C2<T> u2 = {a0.x + c2 * p1.x + c1 * p2.x, a0.y + c2 * p1.y + c1 * p2.y};
  C2<T> v1 = {s1 * m1.x + s2 * m2.x, s1 * m1.y + s2 * m2.y};
  C2<T> v2 = {s2 * m1.x - s1 * m2.x, s2 * m1.y - s1 * m2.y};
  // X_k = u - i v  (forward sign)
  a[1] = {u1.x + v1.y, u1.y - v1.x};
  a[4] = {u1.x - v1.y, u1.y + v1.x};
  a[2] = {u2.x + v2.y, u2.y - v2.x};
  a[3] = {u2.x - v2.y, u2.y + v2.x};
}

template <class T> CB_HD void dft8(C2<T>* a) {
  const T h = T(0.70710678118654752440084436210485);
  dft4(a[0], a[2], a[4], a[6]);  // E_k in a0,a2,a4,a6
  dft4(a[1], a[3], a[5], a[7]);  // O_k in a1,a3,a5,a7
  C2<T> o1 = {h * (a[3].x + a[3].y), h * (a[3].y - a[3].x)};   // w8^1 * O1
  C2<T> o2 = mul_mi(a[5]);                                        // w8^2 * O2
  C2<T> o3 = {h * (a[7].y - a[7].x), -h * (a[7].x + a[7].y)};  // w8^3 * O3
  C2<T> e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
  a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
  a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
  a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
  a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <class T> CB_HD void dft16(C2<T>* a) {
  // n = 4 n1 + n2, k = k1 + 4 k2
  const T c1 = T(0.92387953251128675612818318939679), s1 = T(0.38268343236508977172845998403040);
  const T h = T(0.70710678118654752440084436210485);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int n2 = 0; n2 < 4; ++n2) dft4(a[n2], a[n2 + 4], a[n2 + 8], a[n2 + 12]);
  // a[n2 + 4 k1] = B[n2][k1]; twiddle w16^(n2 k1)
  const C2<T> w1 = {c1, -s1}, w2 = {h, -h}, w3 = {s1, -c1};
  const C2<T> w6 = {-h, -h}, w9 = {-c1, s1};
  a[1 + 4] = cmul(a[1 + 4], w1);  a[1 + 8] = cmul(a[1 + 8], w2);  a[1 + 12] = cmul(a[1 + 12], w3);
  a[2 + 4] = cmul(a[2 + 4], w2);  a[2 + 8] = mul_mi(a[2 + 8]);    a[2 + 12] = cmul(a[2 + 12], w6);
  a[3 + 4] = cmul(a[3 + 4], w3);  a[3 + 8] = cmul(a[3 + 8], w6);  a[3 + 12] = cmul(a[3 + 12], w9);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k1 = 0; k1 < 4; ++k1) dft4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
  // now a[4 k1 + k2] = X[k1 + 4 k2]  -> transpose to natural order
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k1 = 0; k1 < 4; ++k1)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k2 = k1 + 1; k2 < 4; ++k2) {
      C2<T> t = a[4 * k1 + k2];
      a[4 * k1 + k2] = a[4 * k2 + k1];
      a[4 * k2 + k1] = t;
    }
}

// generic O(R^2) butterfly for odd primes not hand-written (7, 11, 13);
// roots taken from the length-M twiddle table: w_R^m = tw[m * (M/R)]
template <class T, int R> CB_HD void dft_generic(C2<T>* a, const C2<T>* tw, int M) {
  C2<T> o[R];
  const int st = M / R;
  for (int r = 0; r < R; ++r) {
    C2<T> acc = a[0];
    for (int q = 1; q < R; ++q) acc = cadd(acc, cmul(a[q], tw[((q * r) % R) * st]));
    o[r] = acc;
  }
  for (int r = 0; r < R; ++r) a[r] = o[r];
}

template <class T, int R> CB_HD void dftR(C2<T>* a, const C2<T>* tw, int M) {
  if (R == 2) dft2(a);
  else if (R == 3) dft3(a);
  else if (R == 4) dft4(a);
  else if (R == 5) dft5(a);
  else if (R == 8) dft8(a);
  else if (R == 16) dft16(a);
  else dft_generic<T, R>(a, tw, M);
}

// ---------------------------------------------------------------------------
// shared-memory layouts.  A "line" holds n reals = M packed complex pairs
// (re at real-index 2j, im at 2j+1).
//   LayX: lanes run ALONG the line (contiguous x transforms); complex index is
//         padded so that strides 1, 8, 64 (radix-8/16 stages and the
//         digit-reversed gather) spread over all banks.
//   LayY: lanes run ACROSS lines (strided y transforms); tile is [index][line],
//         conflict-free for every stride because the line index is fastest.
// ---------------------------------------------------------------------------
CB_HD int padc(int j) { return j + (j >> 3) + (j >> 6); }

struct LayX {
  int ls;  // doubles per line (2*padc(M-1)+2, even)
  CB_HD int at(int c, int j, int h) const { return c * ls + 2 * padc(j) + h; }
  static CB_HD int line_len(int M) { return 2 * (padc(M > 0 ? M - 1 : 0) + 1); }
  // work item -> (line, sub index); sub index fastest across lanes
  static CB_HD void split(int w, int nl, int nsub, int& c, int& b) { (void)nl; c = w / nsub; b = w - c * nsub; }
};

struct LayY {
  int cx;  // lines per tile (fastest)
  CB_HD int at(int c, int j, int h) const { return (2 * j + h) * cx + c; }
  static CB_HD void split(int w, int nl, int nsub, int& c, int& b) { (void)nsub; b = w / nl; c = w - b * nl; }
};

// ---------------------------------------------------------------------------
// plan data visible to kernels
// ---------------------------------------------------------------------------
#define CB_MAX_STAGES 8
template <class T> struct FftDev {
  int n;        // logical real length
  int M;        // n / 2
  int kind;     // Kind
  int nstages;
  int radix[CB_MAX_STAGES];
  const C2<T>* tw;        // M entries  exp(-2 pi i t / M)
  const C2<T>* twp;       // M/2+1 entries exp(-2 pi i k / n)
  const C2<T>* mak;       // M+1 entries (cos, sin)(pi k / (2 n))
  const uint16_t* rev;    // M entries: position of frequency k after the DIF stages
};

// one butterfly of stage `s`: sub-length Ns, Bs = M / Ns blocks
template <class T, int R, class Lay>
CB_HD void stage_item(T* s, const Lay& lay, int c, int b, int Ns, int Bs, const C2<T>* tw, int M) {
  const int L = Ns / R;
  const int blk = b / L;
  const int o = b - blk * L;
  const int base = blk * Ns + o;
  C2<T> a[R];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < R; ++q) {
    const int j = base + q * L;
    a[q].x = s[lay.at(c, j, 0)];
    a[q].y = s[lay.at(c, j, 1)];
  }
  dftR<T, R>(a, tw, M);
  const int tstep = o * Bs;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < R; ++r) {
    C2<T> v = a[r];
    if (r > 0 && L > 1) v = cmul(v, tw[tstep * r]);
    const int j = base + r * L;
    s[lay.at(c, j, 0)] = v.x;
    s[lay.at(c, j, 1)] = v.y;
  }
}

template <class T, class Lay>
CB_HD void stage_item_dyn(int R, T* s, const Lay& lay, int c, int b, int Ns, int Bs, const C2<T>* tw, int M) {
  switch (R) {
    case 16: stage_item<T, 16, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 8: stage_item<T, 8, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 4: stage_item<T, 4, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 2: stage_item<T, 2, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 3: stage_item<T, 3, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 5: stage_item<T, 5, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 7: stage_item<T, 7, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 11: stage_item<T, 11, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    case 13: stage_item<T, 13, Lay>(s, lay, c, b, Ns, Bs, tw, M); break;
    default: break;
  }
}

// index of input sample i inside the real sequence v that is packed as
// complex pairs (Makhoul permutation for the cosine/sine kinds)
CB_HD int vindex(int kind, int n, int i) {
  if (kind == K_R2HC || kind == K_HC2R) return i;
  return (i & 1) ? (n - 1 - (i >> 1)) : (i >> 1);
}

// ---- forward kinds: load one input sample into the tile --------------------
template <class T, class Lay>
CB_HD void fwd_load_item(const FftDev<T>& P, T* s, const Lay& lay, int c, int i, T x) {
  if (P.kind == K_RODFT10 && (i & 1)) x = -x;
  const int v = vindex(P.kind, P.n, i);
  s[lay.at(c, v >> 1, v & 1)] = x;
}

// ---- forward kinds: post-process pair (k, M-k), k = 0..M/2, write outputs ---
// `out(idx, value)` stores output element idx of this line.
template <class T, class Lay, class Out>
CB_HD void fwd_post_item(const FftDev<T>& P, const T* s, const Lay& lay, int c, int k, Out out) {
  const int M = P.M, n = P.n;
  const int pk = P.rev[k];
  C2<T> zk = {s[lay.at(c, pk, 0)], s[lay.at(c, pk, 1)]};
  C2<T> vk, vm;  // V_k, V_{M-k}
  if (k == 0) {
    vk = {zk.x + zk.y, T(0)};  // V_0
    vm = {zk.x - zk.y, T(0)};  // V_M
  } else {
    const int pm = P.rev[M - k];
    C2<T> zm = {s[lay.at(c, pm, 0)], s[lay.at(c, pm, 1)]};
    C2<T> e = {T(0.5) * (zk.x + zm.x), T(0.5) * (zk.y - zm.y)};
    C2<T> o = {T(0.5) * (zk.y + zm.y), T(0.5) * (zm.x - zk.x)};  // -i (zk - conj zm)/2
    C2<T> wo = cmul(P.twp[k], o);
    vk = cadd(e, wo);
    C2<T> d = csub(e, wo);
    vm = {d.x, -d.y};
  }
  const int km = M - k;
  if (P.kind == K_R2HC) {
    if (k == 0) {
      out(0, vk.x);
      out(M, vm.x);
    } else {
      out(k, vk.x);
      out(n - k, vk.y);
      if (km != k) {
        out(km, vm.x);
        out(n - km, vm.y);
      }
    }
    return;
  }
  // Makhoul post-twiddle: X_k = 2 (c V.re + s V.im), X_{n-k} = 2 (s V.re - c V.im)
  const bool rev_out = (P.kind == K_RODFT10);
  {
    C2<T> cs = P.mak[k];
    T xa = T(2) * (cs.x * vk.x + cs.y * vk.y);
    out(rev_out ? n - 1 - k : k, xa);
    if (k > 0) {
      T xb = T(2) * (cs.y * vk.x - cs.x * vk.y);
      out(rev_out ? k - 1 : n - k, xb);
    }
  }
  if (km != k) {  // also true for k == 0 (km == M): V_M is real, X_M = 2 c_M V_M
    C2<T> cs = P.mak[km];
    T xa = T(2) * (cs.x * vm.x + cs.y * vm.y);
    out(rev_out ? n - 1 - km : km, xa);
    if (k > 0) {
      T xb = T(2) * (cs.y * vm.x - cs.x * vm.y);
      out(rev_out ? km - 1 : n - km, xb);
    }
  }
}

// ---- backward kinds: pre-process pair (k, M-k), k = 0..M/2 -----------------
// `in(idx)` returns input element idx of this line (spectrum, FFTW order).
// Writes Z_k and Z_{M-k} (natural positions) with re/im SWAPPED so that the
// forward engine performs the inverse transform.
template <class T, class Lay, class In>
CB_HD void bwd_pre_item(const FftDev<T>& P, T* s, const Lay& lay, int c, int k, In in) {
  const int M = P.M, n = P.n;
  const int km = M - k;
  C2<T> wk, wm;  // W_k, W_{M-k}
  if (P.kind == K_HC2R) {
    wk = {in(k), (k == 0) ? T(0) : in(n - k)};
    wm = {in(km), (km == M) ? T(0) : in(n - km)};
  } else {
    const bool rv = (P.kind == K_RODFT01);
    // X_idx with X_n := 0; reversed input for the sine kind
    T xk = rv ? in(n - 1 - k) : in(k);
    T xnk = (k == 0) ? T(0) : (rv ? in(k - 1) : in(n - k));
    T xm = rv ? in(n - 1 - km) : in(km);
    T xnm = rv ? in(km - 1) : in(n - km);  // km >= 1 always (k <= M/2, M >= 1)
    C2<T> ck = P.mak[k], cm = P.mak[km];
    wk = {ck.x * xk + ck.y * xnk, ck.y * xk - ck.x * xnk};
    wm = {cm.x * xm + cm.y * xnm, cm.y * xm - cm.x * xnm};
  }
  // A_k = W_k + conj(W_{M-k});  B_k = (W_k - conj(W_{M-k})) * conj(twp[k]);  Z_k = A_k + i B_k
  C2<T> a = {wk.x + wm.x, wk.y - wm.y};
  C2<T> d = {wk.x - wm.x, wk.y + wm.y};
  C2<T> t = P.twp[k];
  C2<T> b = {d.x * t.x + d.y * t.y, d.y * t.x - d.x * t.y};  // d * conj(t)
  C2<T> zk = {a.x - b.y, a.y + b.x};
  // swapped store: re -> slot 1, im -> slot 0
  s[lay.at(c, k % M, 1)] = zk.x;
  s[lay.at(c, k % M, 0)] = zk.y;
  if (km != k && km != M) {
    // Z_{M-k} = conj(A_k) + i conj(B_k)
    C2<T> zm = {a.x + b.y, -a.y + b.x};
    s[lay.at(c, km, 1)] = zm.x;
    s[lay.at(c, km, 0)] = zm.y;
  }
}

// ---- backward kinds: fetch output sample i after the stages ----------------
template <class T, class Lay>
CB_HD T bwd_out_item(const FftDev<T>& P, const T* s, const Lay& lay, int c, int i) {
  const int v = vindex(P.kind, P.n, i);
  const int j = v >> 1, h = v & 1;
  const int pj = P.rev[j];
  T x = s[lay.at(c, pj, 1 - h)];  // swapped read
  if (P.kind == K_RODFT01 && (i & 1)) x = -x;
  return x;
}

}  // namespace cb
