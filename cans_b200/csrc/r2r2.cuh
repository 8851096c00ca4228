// Two-for-one register-resident batched r2r transforms (sm_100a fast path).
//
// Replaces one `call fft(arrplan(idir,fb), arr)` of the reference CPU path
// (/root/reference/src/fft.f90:247-258; FFTW guru r2r plans :83-97,:148-162)
// and the cuFFT + OpenACC Makhoul passes of the reference GPU path
// (src/fft.f90:369-401, :452-767) with ONE kernel per transform that touches
// HBM once (read the field, write the field).
//
// Algorithm.  Two real lines a, b are transformed by one complex FFT of the
// full length N on z = a + i b ("two-for-one"); the spectra are separated with
//     A_k = (Z_k + conj Z_{N-k}) / 2,   B_k = (Z_k - conj Z_{N-k}) / (2i)
// which needs no twiddle factors.  In x (contiguous lines) the pair is two
// consecutive lines; in y (strided lines) it is two adjacent x columns, so one
// 16-byte element of a 128-byte row is one complex sample.  The cosine / sine
// kinds use Makhoul's permutation v_j = x_{2j} | x_{2(N-1-j)+1} at load time
// and the quarter-wave rotation in the separation pass.
//
// Forward kinds run decimation-in-frequency: natural-order samples go straight
// from global memory into registers (coalesced), each thread owns E = N / TPL
// samples, stages exchange through shared memory in place, the spectrum ends
// digit-reversed in shared memory and the separation pass gathers (k, N-k).
// Backward kinds run the transposed network (decimation-in-time): the
// pre-pass scatters Z_k to its digit-reversed slot (re/im swapped, which turns
// the forward butterflies into the inverse transform), and the last stage
// leaves natural-order samples in registers that are stored coalesced.
//
// Shared memory: x mode keeps one region of N complex per transform with an
// XOR swizzle of the low 3 (FP64) / 4 (FP32) index bits, which makes the
// strided butterfly accesses of every stage conflict free; y mode keeps
// [position][column pair], whose rows are exactly 128 bytes, so any set of
// rows is conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "fft_engine.cuh"  // Kind constants
#include "r2_fill.cuh"    // R2Fill: the fused fillps (+ updt_rhs_b) source of the forward x transform

namespace cb {

template <class T> struct __align__(sizeof(T) * 2) Cx { T x, y; };

template <class T> __device__ __forceinline__ Cx<T> cx_add(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <class T> __device__ __forceinline__ Cx<T> cx_sub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <class T> __device__ __forceinline__ Cx<T> cx_mul(Cx<T> a, Cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <class T> __device__ __forceinline__ Cx<T> cx_mi(Cx<T> a) { return {a.y, -a.x}; }  // * (-i)
__device__ __forceinline__ Cx<double> cx_ldg(const Cx<double>* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}
__device__ __forceinline__ Cx<float> cx_ldg(const Cx<float>* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return {v.x, v.y};
}

// field loads that bypass L1 allocation (the field is streamed once; L1 is left to the twiddle tables) and
// streaming (evict-first) stores
__device__ __forceinline__ double ld_na(const double* p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_na(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ Cx<double> ld_na(const Cx<double>* p) {
  Cx<double> v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ Cx<float> ld_na(const Cx<float>* p) {
  Cx<float> v;
  asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_cs(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(Cx<double>* p, Cx<double> v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }
__device__ __forceinline__ void st_cs(Cx<float>* p, Cx<float> v) { __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }
#define CB_R2_LD_NA 1   // R2Args::flags
#define CB_R2_ST_CS 2
// x transforms of a periodic direction inside the solve: the spectrum is kept in SPLIT order
//   (r0, r1, .., r[n/2-1] | r[n/2], i1, .., i[n/2-1])      instead of FFTW's halfcomplex (r0, .., r[n/2], i[n/2-1], .., i1),
// i.e. the imaginary part of mode k sits n/2 positions after its real part.  Nothing between the two x transforms cares
// about the order (the eigenvalues are permuted to match), and the tridiagonal stage can then share the pivots of whole
// 128-byte tiles between the two halves (ThomasDev::dx).
#define CB_R2_XSPLIT 4

// ---- register butterflies: forward sign, natural in, natural out --------------------------------
template <class T> __device__ __forceinline__ void bf2(Cx<T>* a) {
  const Cx<T> t = a[0];
  a[0] = cx_add(t, a[1]);
  a[1] = cx_sub(t, a[1]);
}
template <class T> __device__ __forceinline__ void bf3(Cx<T>* a) {
  const T c = T(0.86602540378443864676372317075294);
  const Cx<T> s = cx_add(a[1], a[2]), d = cx_sub(a[1], a[2]);
  const Cx<T> m = {a[0].x - T(0.5) * s.x, a[0].y - T(0.5) * s.y};
  a[0] = cx_add(a[0], s);
  a[1] = {m.x + c * d.y, m.y - c * d.x};
  a[2] = {m.x - c * d.y, m.y + c * d.x};
}
template <class T> __device__ __forceinline__ void bf4(Cx<T>& a0, Cx<T>& a1, Cx<T>& a2, Cx<T>& a3) {
  const Cx<T> t0 = cx_add(a0, a2), t1 = cx_sub(a0, a2), t2 = cx_add(a1, a3), t3 = cx_sub(a1, a3);
  a0 = cx_add(t0, t2);
  a2 = cx_sub(t0, t2);
  a1 = {t1.x + t3.y, t1.y - t3.x};
  a3 = {t1.x - t3.y, t1.y + t3.x};
}
template <class T> __device__ __forceinline__ void bf5(Cx<T>* a) {
  const T c1 = T(0.30901699437494742410229341718282), c2 = T(-0.80901699437494742410229341718282);
  const T s1 = T(0.95105651629515357211643933337938), s2 = T(0.58778525229247312916870595463907);
  const Cx<T> p1 = cx_add(a[1], a[4]), m1 = cx_sub(a[1], a[4]);
  const Cx<T> p2 = cx_add(a[2], a[3]), m2 = cx_sub(a[2], a[3]);
  const Cx<T> a0 = a[0];
  a[0] = {a0.x + p1.x + p2.x, a0.y + p1.y + p2.y};
  const Cx<T> u1 = {a0.x + c1 * p1.x + c2 * p2.x, a0.y + c1 * p1.y + c2 * p2.y};
  const Cx<T> u2 = {a0.x + c2 * p1.x + c1 * p2.x, a0.y + c2 * p1.y + c1 * p2.y};
  const Cx<T> v1 = {s1 * m1.x + s2 * m2.x, s1 * m1.y + s2 * m2.y};
  const Cx<T> v2 = {s2 * m1.x - s1 * m2.x, s2 * m1.y - s1 * m2.y};
  a[1] = {u1.x + v1.y, u1.y - v1.x};
  a[4] = {u1.x - v1.y, u1.y + v1.x};
  a[2] = {u2.x + v2.y, u2.y - v2.x};
  a[3] = {u2.x - v2.y, u2.y + v2.x};
}
template <class T> __device__ __forceinline__ void bf8(Cx<T>* a) {
  const T h = T(0.70710678118654752440084436210485);
  bf4(a[0], a[2], a[4], a[6]);
  bf4(a[1], a[3], a[5], a[7]);
  const Cx<T> o1 = {h * (a[3].x + a[3].y), h * (a[3].y - a[3].x)};
  const Cx<T> o2 = cx_mi(a[5]);
  const Cx<T> o3 = {h * (a[7].y - a[7].x), -h * (a[7].x + a[7].y)};
  const Cx<T> e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
  a[0] = cx_add(e0, o0); a[4] = cx_sub(e0, o0);
  a[1] = cx_add(e1, o1); a[5] = cx_sub(e1, o1);
  a[2] = cx_add(e2, o2); a[6] = cx_sub(e2, o2);
  a[3] = cx_add(e3, o3); a[7] = cx_sub(e3, o3);
}
// 12 = 4 x 3: n = 3 n1 + n2, k = k1 + 4 k2
template <class T> __device__ __forceinline__ void bf12(Cx<T>* a) {
  const T c = T(0.86602540378443864676372317075294);
#pragma unroll
  for (int n2 = 0; n2 < 3; ++n2) bf4(a[n2], a[n2 + 3], a[n2 + 6], a[n2 + 9]);
  // a[n2 + 3 k1] = B[n2][k1]; twiddle w12^(n2 k1)
  const Cx<T> w1 = {c, T(-0.5)}, w2 = {T(0.5), -c}, w4 = {T(-0.5), -c};
  a[1 + 3] = cx_mul(a[1 + 3], w1); a[1 + 6] = cx_mul(a[1 + 6], w2); a[1 + 9] = cx_mi(a[1 + 9]);
  a[2 + 3] = cx_mul(a[2 + 3], w2); a[2 + 6] = cx_mul(a[2 + 6], w4); a[2 + 9] = {-a[2 + 9].x, -a[2 + 9].y};
  Cx<T> o[12];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    Cx<T> b[3] = {a[3 * k1], a[3 * k1 + 1], a[3 * k1 + 2]};
    bf3(b);
    o[k1] = b[0]; o[k1 + 4] = b[1]; o[k1 + 8] = b[2];
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) a[k] = o[k];
}
// 6 = 2 x 3: n = 3 n1 + n2, k = k1 + 2 k2
template <class T> __device__ __forceinline__ void bf6(Cx<T>* a) {
  const T c = T(0.86602540378443864676372317075294);
#pragma unroll
  for (int n2 = 0; n2 < 3; ++n2) {
    const Cx<T> t = a[n2];
    a[n2] = cx_add(t, a[n2 + 3]);
    a[n2 + 3] = cx_sub(t, a[n2 + 3]);
  }
  const Cx<T> w1 = {T(0.5), -c}, w2 = {T(-0.5), -c};
  a[1 + 3] = cx_mul(a[1 + 3], w1);
  a[2 + 3] = cx_mul(a[2 + 3], w2);
  Cx<T> o[6];
#pragma unroll
  for (int k1 = 0; k1 < 2; ++k1) {
    Cx<T> b[3] = {a[3 * k1], a[3 * k1 + 1], a[3 * k1 + 2]};
    bf3(b);
    o[k1] = b[0]; o[k1 + 2] = b[1]; o[k1 + 4] = b[2];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) a[k] = o[k];
}
template <class T> __device__ __forceinline__ void bf16(Cx<T>* a) {
  const T c1 = T(0.92387953251128675612818318939679), s1 = T(0.38268343236508977172845998403040);
  const T h = T(0.70710678118654752440084436210485);
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) bf4(a[n2], a[n2 + 4], a[n2 + 8], a[n2 + 12]);
  const Cx<T> w1 = {c1, -s1}, w2 = {h, -h}, w3 = {s1, -c1}, w6 = {-h, -h}, w9 = {-c1, s1};
  a[5] = cx_mul(a[5], w1);   a[9] = cx_mul(a[9], w2);   a[13] = cx_mul(a[13], w3);
  a[6] = cx_mul(a[6], w2);   a[10] = cx_mi(a[10]);      a[14] = cx_mul(a[14], w6);
  a[7] = cx_mul(a[7], w3);   a[11] = cx_mul(a[11], w6); a[15] = cx_mul(a[15], w9);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) bf4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = k1 + 1; k2 < 4; ++k2) {
      const Cx<T> t = a[4 * k1 + k2];
      a[4 * k1 + k2] = a[4 * k2 + k1];
      a[4 * k2 + k1] = t;
    }
}
template <class T, int R> __device__ __forceinline__ void bfR(Cx<T>* a) {
  static_assert(R == 1 || R == 2 || R == 3 || R == 4 || R == 5 || R == 6 || R == 8 || R == 12 || R == 16, "radix");
  if (R == 2) bf2(a);
  else if (R == 3) bf3(a);
  else if (R == 4) bf4(a[0], a[1], a[2], a[3]);
  else if (R == 5) bf5(a);
  else if (R == 6) bf6(a);
  else if (R == 8) bf8(a);
  else if (R == 12) bf12(a);
  else if (R == 16) bf16(a);
}

// ---- compile-time plan --------------------------------------------------------------------------
// N = R0 R1 R2 R3 complex length, TPL threads per transform (E = N / TPL samples per thread, every
// radix divides E), G = transforms per CTA (x mode) / column pairs per CTA (y mode), MINB = CTAs per
// SM the register allocation must allow (__launch_bounds__).
template <int N_, int TPL_, int G_, int MINB_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1> struct R2Cfg {
  static constexpr int N = N_, TPL = TPL_, G = G_, E = N_ / TPL_, MINB = MINB_;
  static constexpr int NS = R3_ > 1 ? 4 : (R2_ > 1 ? 3 : (R1_ > 1 ? 2 : 1));
  __host__ __device__ static constexpr int R(int s) { return s == 0 ? R0_ : (s == 1 ? R1_ : (s == 2 ? R2_ : R3_)); }
  __host__ __device__ static constexpr int Ns(int s) { return s == 0 ? N_ : (s == 1 ? N_ / R0_ : (s == 2 ? N_ / (R0_ * R1_) : N_ / (R0_ * R1_ * R2_))); }
  static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
  static_assert(N_ % TPL_ == 0 && E % R0_ == 0 && E % R1_ == 0 && E % R2_ == 0 && E % R3_ == 0, "every radix must divide E");
  // slot of frequency k after the DIF stages (digit reversal)
  __host__ __device__ static constexpr int rev(int k) {
    int pos = 0;
    pos += (k % R0_) * (N_ / R0_); k /= R0_;
    pos += (k % R1_) * (N_ / (R0_ * R1_)); k /= R1_;
    pos += (k % R2_) * (N_ / (R0_ * R1_ * R2_)); k /= R2_;
    pos += k;
    return pos;
  }
};

// distributed y transforms: row i of a line lives at ptr + g * gs + x on some GPU of the box (peer-mapped)
template <class T> struct __align__(16) R2Row { T* ptr; long long gs; };

template <class T> struct R2Args {
  const T* in;
  T* out;
  long long in_es, out_es;   // element stride along the transform (1 in x mode, the row pitch in y mode)
  long long in_ls, out_ls;   // stride between consecutive lines of a group
  long long in_gs, out_gs;   // stride between groups (z planes)
  int lines_per_group, ngroups, line_len;  // line_len >= N: the tail is copied, not transformed
  int kind;
  const Cx<T>* tw[4];        // per stage s: (R_s - 1) * L_s entries, w_{Ns}^{o r} at [(r-1) L + o]
  const Cx<T>* mak;          // (cos, sin)(pi k / (2N)), k = 0..N/2
  const R2Row<T>* row_tab;   // SPLIT kernels: where each output (forward) / input (backward) row lives
  int flags;                 // CB_R2_LD_NA | CB_R2_ST_CS
  int smem_pad = 0;          // host only: extra dynamic shared memory per CTA, to cap the CTAs per SM of a kernel that shares the GPU
  int x0 = 0, g0 = 0;        // SPLIT kernels launched on a window of the slab: first column / first plane of the window
                             // (the row table addresses whole slab rows; in / out already point at the window)
};

template <class T, class Cfg, bool YMODE> struct R2Lay {
  static constexpr int SW = sizeof(T) == 8 ? 3 : 4;
  static constexpr int NP = (Cfg::N + (1 << SW) - 1) & ~((1 << SW) - 1);
  static constexpr size_t smem_bytes() { return (YMODE ? (size_t)Cfg::N * Cfg::G : (size_t)NP * Cfg::G) * sizeof(Cx<T>); }
  // encoded position: x mode = XOR-swizzled index (linear over GF(2)), y mode = the index itself
  __host__ __device__ static constexpr int enc(int pos) {
    return YMODE ? pos : (pos ^ (((pos >> SW) ^ (pos >> (2 * SW))) & ((1 << SW) - 1)));
  }
  __device__ static __forceinline__ int at(int c, int pos) {
    if (YMODE) return pos * Cfg::G + c;
    return c * NP + enc(pos);
  }
  // slot of position pa + pb given enc(pa), enc(pb), for pa and pb with DISJOINT bit sets (power-of-two plans: a
  // per-thread part and a compile-time part), so that enc(pa + pb) = enc(pa) ^ enc(pb) (x) / enc(pa) + enc(pb) (y)
  __device__ static __forceinline__ int join(int c, int ea, int eb) {
    if (YMODE) return (ea + eb) * Cfg::G + c;
    return c * NP + (ea ^ eb);
  }
};

template <int ID, int CNT> __device__ __forceinline__ void r2_named_bar() {
#if defined(CB_EMU_HOST)   // tests/emu/emu_r2r2.cpp: the kernels' source under g++, CUDA threads as host threads
  cb_emu_named_bar(ID, CNT);
#else
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(CNT) : "memory");
#endif
}
// barrier among the TPL threads of one transform.  Barrier ids are compile-time constants so that
// ptxas reserves G + 1 hardware barriers per CTA, not all 16 (which would cap the CTAs per SM at 4).
// does r2_group_sync involve only the threads of one transform (so that a whole transform may exit early)?
template <class Cfg> __host__ __device__ constexpr bool r2_group_local() {
  return Cfg::G == 1 || (Cfg::TPL <= 32 && (32 % Cfg::TPL) == 0) || ((Cfg::TPL % 32) == 0 && Cfg::G <= 8);
}
template <class Cfg, bool YMODE> __device__ __forceinline__ void r2_group_sync(int c) {
  if (YMODE || Cfg::G == 1) __syncthreads();
  else if (Cfg::TPL <= 32 && (32 % Cfg::TPL) == 0) __syncwarp();
  else if ((Cfg::TPL % 32) == 0 && Cfg::G <= 8) {
    switch (c) {
      case 0: r2_named_bar<1, Cfg::TPL>(); break;
      case 1: r2_named_bar<2, Cfg::TPL>(); break;
      case 2: r2_named_bar<3, Cfg::TPL>(); break;
      case 3: r2_named_bar<4, Cfg::TPL>(); break;
      case 4: r2_named_bar<5, Cfg::TPL>(); break;
      case 5: r2_named_bar<6, Cfg::TPL>(); break;
      case 6: r2_named_bar<7, Cfg::TPL>(); break;
      default: r2_named_bar<8, Cfg::TPL>(); break;
    }
  } else __syncthreads();
}

// where my transform lives in global memory
template <class T> struct R2Loc {
  long long ia, ib, oa, ob;  // offsets of sequences a and b (x mode: two lines; y mode: ib = ia + 1)
  bool has_a, has_b;
  int g, x;                  // y mode: group (z plane) and first column of my pair
};

template <class T, class Cfg, bool YMODE>
__device__ __forceinline__ R2Loc<T> r2_locate(const R2Args<T>& A, int c) {
  R2Loc<T> L;
  if (YMODE) {
    const int tiles = (A.lines_per_group + 2 * Cfg::G - 1) / (2 * Cfg::G);
    const int g = blockIdx.x / tiles;
    const int x = (blockIdx.x - g * tiles) * 2 * Cfg::G + 2 * c;
    L.ia = (long long)g * A.in_gs + (long long)x * A.in_ls;
    L.oa = (long long)g * A.out_gs + (long long)x * A.out_ls;
    L.ib = L.ia + A.in_ls;
    L.ob = L.oa + A.out_ls;
    L.has_a = x < A.lines_per_group;
    L.has_b = x + 1 < A.lines_per_group;
    L.g = g + A.g0;
    L.x = x + A.x0;
  } else {
    // 32-bit index arithmetic (the launcher checks that the line count fits)
    const unsigned lpg = (unsigned)A.lines_per_group;
    const unsigned nl = lpg * (unsigned)A.ngroups;
    const unsigned la = (blockIdx.x * (unsigned)Cfg::G + (unsigned)c) * 2u, lb = la + 1u;
    const unsigned ga = la / lpg, ja = la - ga * lpg;
    const unsigned gb = (ja + 1u == lpg) ? ga + 1u : ga, jb = (ja + 1u == lpg) ? 0u : ja + 1u;
    L.ia = (long long)ga * A.in_gs + (long long)ja * A.in_ls;
    L.oa = (long long)ga * A.out_gs + (long long)ja * A.out_ls;
    L.ib = (long long)gb * A.in_gs + (long long)jb * A.in_ls;
    L.ob = (long long)gb * A.out_gs + (long long)jb * A.out_ls;
    L.has_a = la < nl;
    L.has_b = lb < nl;
    if (!L.has_b) { L.ib = L.ia; L.ob = L.oa; }   // odd line count: b aliases a, and a is stored last (r2_store)
    L.g = 0;
    L.x = 0;
  }
  return L;
}

// split (peer-mapped) rows of the distributed y transforms
__device__ __forceinline__ R2Row<double> r2_row(const R2Row<double>* tab, int i) {
  const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(tab + i));
  R2Row<double> r;
  r.ptr = reinterpret_cast<double*>(v.x);
  r.gs = v.y;
  return r;
}
__device__ __forceinline__ R2Row<float> r2_row(const R2Row<float>* tab, int i) {
  const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(tab + i));
  R2Row<float> r;
  r.ptr = reinterpret_cast<float*>(v.x);
  r.gs = v.y;
  return r;
}
// Makhoul permutation: position j of v holds input sample vperm(j)
__device__ __forceinline__ int r2_vperm(bool trig, int n, int j) {
  if (!trig) return j;
  const int h = (n + 1) >> 1;
  return j < h ? 2 * j : 2 * (n - 1 - j) + 1;
}

// Row accessor of one transform: rows `base + off` with a per-thread `base` folded into the pointers once and a
// compile-time `off`, so that the unrolled loads / stores use immediate offsets (x mode) or one multiply-add
// (y mode).  x mode: two lines a, b; y mode: one 16-byte column pair per row; SPLIT: rows located through the
// peer-mapped row table.
template <class T, bool YMODE, bool SPLIT, bool FULL> struct R2Rows {
  T* pa;
  T* pb;
  long long es;
  const R2Row<T>* tab;
  int g, x;
  bool has_a, has_b;
  __device__ __forceinline__ Cx<T> load(int off) const {
    Cx<T> v;
    if (SPLIT) {
      const R2Row<T> e = r2_row(tab, off);
      const T* q = e.ptr + (long long)g * e.gs + x;
      if (FULL || has_b) return *reinterpret_cast<const Cx<T>*>(q);
      v.x = has_a ? *q : T(0);
      v.y = T(0);
    } else if (YMODE) {
      const T* q = pa + (long long)off * es;
      if (FULL || has_b) return *reinterpret_cast<const Cx<T>*>(q);
      v.x = has_a ? *q : T(0);
      v.y = T(0);
    } else {
      v.x = pa[off];
      v.y = pb[off];
    }
    return v;
  }
  __device__ __forceinline__ void store(int off, T a, T b) const {
    if (SPLIT) {
      const R2Row<T> e = r2_row(tab, off);
      T* q = e.ptr + (long long)g * e.gs + x;
      if (FULL || has_b) { *reinterpret_cast<Cx<T>*>(q) = Cx<T>{a, b}; return; }
      if (has_a) *q = a;
    } else if (YMODE) {
      T* q = pa + (long long)off * es;
      if (FULL || has_b) { *reinterpret_cast<Cx<T>*>(q) = Cx<T>{a, b}; return; }
      if (has_a) *q = a;
    } else {
      pb[off] = b;   // b first: when the second line is missing it aliases the first and a must win
      pa[off] = a;
    }
  }
};
template <class T, bool YMODE, bool SPLIT, bool FULL>
__device__ __forceinline__ R2Rows<T, YMODE, SPLIT, FULL> r2_rows(const T* p, long long oa, long long ob, long long es,
                                                                  const R2Row<T>* tab, const R2Loc<T>& L, int base) {
  R2Rows<T, YMODE, SPLIT, FULL> R;
  T* q = const_cast<T*>(p);
  R.pa = q + oa + (YMODE ? (long long)base * es : (long long)base);
  R.pb = q + ob + (long long)base;
  R.es = es;
  R.tab = SPLIT ? tab + base : nullptr;
  R.g = L.g; R.x = L.x; R.has_a = L.has_a; R.has_b = L.has_b;
  return R;
}

// kernel arguments of the fused variant: the transform's own arguments followed by the source
template <class T> struct R2ArgsFill : R2Args<T> { R2Fill<T> F; };

// thread part of the positions a thread touches in stage S (the part that does not depend on the unrolled
// indices m, q), and the compile-time remainder; their bit sets are disjoint for power-of-two plans
template <class Cfg, int S> __device__ __forceinline__ int r2_stage_tpos(int t) {
  constexpr int Rs = Cfg::R(S), Ns = Cfg::Ns(S), L = Ns / Rs;
  return (L >= Cfg::TPL) ? t : (t / L) * Ns + (t % L);
}
template <class Cfg, int S> __host__ __device__ constexpr int r2_stage_cpos(int m, int q) {
  constexpr int Rs = Cfg::R(S), Ns = Cfg::Ns(S), L = Ns / Rs;
  return ((Cfg::TPL * m) / L) * Ns + ((Cfg::TPL * m) % L) + q * L;
}

// a[r] *= w^(o r), r = 1 .. Rs-1, with w^(o r) at tab[(r-1) L].  The table is read for r = 1 and the even r only;
// the odd ones are w^(o (r-1)) w^o (one extra rounding, ~1e-16): the twiddle loads were 16 % of all L1 wavefronts
// of the x kernels (ncu, profiles/r1g_stage_kernels_ncu.md), this halves them at the cost of 4 flops each.
template <class T, int Rs>
__device__ __forceinline__ void r2_apply_twiddles(Cx<T>* a, const Cx<T>* tab, int L) {
  const Cx<T> w1 = cx_ldg(tab);
  a[1] = cx_mul(a[1], w1);
#pragma unroll
  for (int r = 2; r < Rs; r += 2) {
    const Cx<T> we = cx_ldg(tab + (r - 1) * L);
    a[r] = cx_mul(a[r], we);
    if (r + 1 < Rs) a[r + 1] = cx_mul(a[r + 1], cx_mul(we, w1));
  }
}

// one DIF stage on the shared tile (S > 0) or on registers already loaded (S == 0)
template <class T, class Cfg, bool YMODE, int S>
__device__ __forceinline__ void r2_dif_stage(Cx<T>* v, Cx<T>* sm, const R2Args<T>& A, int c, int t) {
  using Lay = R2Lay<T, Cfg, YMODE>;
  constexpr int Rs = Cfg::R(S), Ns = Cfg::Ns(S), L = Ns / Rs, NB = Cfg::E / Rs, TPL = Cfg::TPL;
  // the slot splits into a per-thread and a compile-time part: always for power-of-two plans (disjoint bit
  // fields), and in y mode (plain index, no swizzle) whenever L and TPL divide one another
  constexpr bool POW2 = ((Cfg::N & (Cfg::N - 1)) == 0) || (YMODE && (L >= TPL ? L % TPL == 0 : TPL % L == 0));
  const int et = Lay::enc(r2_stage_tpos<Cfg, S>(t));
#pragma unroll
  for (int m = 0; m < NB; ++m) {
    const int u = t + TPL * m;
    const int blk = u / L, o = u - blk * L;
    const int base = blk * Ns + o;
    Cx<T>* a = (S == 0) ? v + m * Rs : v;
    if (S > 0) {
#pragma unroll
      for (int q = 0; q < Rs; ++q)
        a[q] = POW2 ? sm[Lay::join(c, et, Lay::enc(r2_stage_cpos<Cfg, S>(m, q)))] : sm[Lay::at(c, base + q * L)];
    }
    bfR<T, Rs>(a);
    if (L > 1) r2_apply_twiddles<T, Rs>(a, A.tw[S] + o, L);
#pragma unroll
    for (int r = 0; r < Rs; ++r) {
      if (POW2) sm[Lay::join(c, et, Lay::enc(r2_stage_cpos<Cfg, S>(m, r)))] = a[r];
      else sm[Lay::at(c, base + r * L)] = a[r];
    }
  }
}

// one DIT stage: twiddle before the butterfly; S == 0 leaves the results in v[]
template <class T, class Cfg, bool YMODE, int S>
__device__ __forceinline__ void r2_dit_stage(Cx<T>* v, Cx<T>* sm, const R2Args<T>& A, int c, int t) {
  using Lay = R2Lay<T, Cfg, YMODE>;
  constexpr int Rs = Cfg::R(S), Ns = Cfg::Ns(S), L = Ns / Rs, NB = Cfg::E / Rs, TPL = Cfg::TPL;
  constexpr bool POW2 = ((Cfg::N & (Cfg::N - 1)) == 0) || (YMODE && (L >= TPL ? L % TPL == 0 : TPL % L == 0));
  const int et = Lay::enc(r2_stage_tpos<Cfg, S>(t));
#pragma unroll
  for (int m = 0; m < NB; ++m) {
    const int u = t + TPL * m;
    const int blk = u / L, o = u - blk * L;
    const int base = blk * Ns + o;
    Cx<T>* a = (S == 0) ? v + m * Rs : v;
#pragma unroll
    for (int q = 0; q < Rs; ++q)
      a[q] = POW2 ? sm[Lay::join(c, et, Lay::enc(r2_stage_cpos<Cfg, S>(m, q)))] : sm[Lay::at(c, base + q * L)];
    if (L > 1) r2_apply_twiddles<T, Rs>(a, A.tw[S] + o, L);
    bfR<T, Rs>(a);
    if (S > 0) {
#pragma unroll
      for (int r = 0; r < Rs; ++r) {
        if (POW2) sm[Lay::join(c, et, Lay::enc(r2_stage_cpos<Cfg, S>(m, r)))] = a[r];
        else sm[Lay::at(c, base + r * L)] = a[r];
      }
    }
  }
}

// Pair pass bookkeeping.  Thread t owns the frequencies k = t + kc, kc = m TPL < N/2, and their mirrors
//   N - k = (TPL - t) + (N - TPL - kc)            (t > 0)
//         = N - kc                                 (t = 0; 0 when kc = 0)
// so global rows are `t + const` ("up") or `(TPL - t) + const` ("dn"), and for power-of-two plans (digit
// reversal = a bit permutation) the shared-memory slots split the same way into a per-thread and a
// compile-time part.
template <class T, class Cfg, bool YMODE> struct R2Pair {
  using Lay = R2Lay<T, Cfg, YMODE>;
  static constexpr bool POW2 = (Cfg::N & (Cfg::N - 1)) == 0;
  int t, c, e_up, e_dn;
  __device__ __forceinline__ R2Pair(int t_, int c_) : t(t_), c(c_) {
    e_up = Lay::enc(Cfg::rev(t_));
    e_dn = Lay::enc(Cfg::rev(t_ ? Cfg::TPL - t_ : 0));
  }
  // slot of Z_k, k = t + kc
  __device__ __forceinline__ int slot_k(int kc) const {
    if (POW2) return Lay::join(c, e_up, Lay::enc(Cfg::rev(kc)));
    return Lay::at(c, Cfg::rev(t + kc));
  }
  // slot of Z_{N-k} (Z_0 for k = 0)
  __device__ __forceinline__ int slot_m(int kc) const {
    if (POW2) {
      const int c1 = Lay::enc(Cfg::rev(Cfg::N - Cfg::TPL - kc)), c0 = Lay::enc(Cfg::rev((Cfg::N - kc) % Cfg::N));
      return Lay::join(c, e_dn, t ? c1 : c0);
    }
    const int k = t + kc;
    return Lay::at(c, Cfg::rev(k ? Cfg::N - k : 0));
  }
};

// ---- forward kinds: R2HC, REDFT10, RODFT10 ---------------------------------------------------------
// ARGS = R2Args<T>: samples come from A.in;  ARGS = R2ArgsFill<T>: x mode only, samples are evaluated from u, v, w (fused
// fillps, A.F).  Five row pointers per line come on top of the samples there: a quarter fewer CTAs per SM than the plain
// kernel keeps the register allocation free of spills (ptxas -v).
template <class ARGS, class T> __host__ __device__ constexpr bool r2_fused() { return !std::is_same<ARGS, R2Args<T>>::value; }
template <class Cfg, bool FUSED> __host__ __device__ constexpr int r2_fwd_minb() {
  return FUSED ? Cfg::MINB - (Cfg::MINB >= 4 ? Cfg::MINB / 4 : 0) : Cfg::MINB;
}
template <class T, class Cfg, bool YMODE, bool SPLIT, int KIND, bool FULL, class ARGS = R2Args<T>>
__global__ void __launch_bounds__(Cfg::TPL* Cfg::G, r2_fwd_minb<Cfg, r2_fused<ARGS, T>()>()) r2r2_fwd_kernel(const ARGS A) {
  static_assert(KIND == K_R2HC || KIND == K_REDFT10 || KIND == K_RODFT10, "forward kinds");
  constexpr bool FUSED = r2_fused<ARGS, T>();
  static_assert(!FUSED || (!YMODE && !SPLIT && FULL), "the fused source feeds the contiguous (x) transforms");
  static_assert(YMODE || r2_group_local<Cfg>(), "x mode relies on per-transform barriers");
  static_assert(!SPLIT || YMODE, "split rows exist only for the strided (y) transforms");
  using C = Cx<T>;
  using Lay = R2Lay<T, Cfg, YMODE>;
  extern __shared__ __align__(16) unsigned char cb_smem_raw[];
  C* sm = reinterpret_cast<C*>(cb_smem_raw);
  constexpr int N = Cfg::N, TPL = Cfg::TPL, E = Cfg::E, G = Cfg::G, NS = Cfg::NS;
  static_assert((TPL & (TPL - 1)) == 0, "threads per transform must be a power of two");
  const int tid = threadIdx.x;
  const int c = YMODE ? tid % G : tid / TPL;
  const int t = (YMODE ? tid / G : tid) & (TPL - 1);
  const R2Loc<T> loc = r2_locate<T, Cfg, YMODE>(A, c);
  if (!YMODE && !loc.has_a) return;   // x mode: group barriers only (r2_group_sync), so a lineless transform may leave
  constexpr bool trig = KIND != K_R2HC;
  constexpr bool neg_odd = KIND == K_RODFT10;
  using RowsIn = R2Rows<T, YMODE, false, FULL>;
  using RowsOut = R2Rows<T, YMODE, SPLIT, FULL>;

  C v[E];
  if constexpr (FUSED) {
    // the same samples, evaluated from the velocity field (R2Fill); x index of a sample = base + compile-time offset
    constexpr int R0 = Cfg::R(0), L0 = N / R0;
    const R2Fill<T>& S = A.F;
    // (plane, row) of my two lines, as r2_locate finds them
    const unsigned lpg = (unsigned)A.lines_per_group;
    const unsigned l0 = (blockIdx.x * (unsigned)G + (unsigned)c) * 2u;
    const unsigned ga = l0 / lpg, ja = l0 - ga * lpg;
    const bool wrap = ja + 1u == lpg;
    const unsigned gb = !loc.has_b ? ga : (wrap ? ga + 1u : ga), jb = !loc.has_b ? ja : (wrap ? 0u : ja + 1u);
    const R2FillLine<T> la(S, loc.ia, (int)ga, (int)ja), lb(S, loc.ib, (int)gb, (int)jb);
    const int b_up = trig ? 2 * t : t, b_dn = -2 * t;
#pragma unroll
    for (int m = 0; m < E / R0; ++m)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int cj = q * L0 + TPL * m;
        C x;
        if (!trig) x = C{la.at(S, b_up + cj), lb.at(S, b_up + cj)};
        else if (cj < N / 2) x = C{la.at(S, b_up + 2 * cj), lb.at(S, b_up + 2 * cj)};
        else {
          x = C{la.at(S, b_dn + 2 * N - 1 - 2 * cj), lb.at(S, b_dn + 2 * N - 1 - 2 * cj)};
          if (neg_odd) x = {-x.x, -x.y};   // odd input rows
        }
        v[m * R0 + q] = x;
      }
  } else {
    // natural-order samples j = t + cj; Makhoul: input row 2 j (cj < N/2) or 2 (N - 1 - j) + 1
    constexpr int R0 = Cfg::R(0), L0 = N / R0;
    const RowsIn in_up = r2_rows<T, YMODE, false, FULL>(A.in, loc.ia, loc.ib, A.in_es, nullptr, loc, trig ? 2 * t : t);
    const RowsIn in_dn = r2_rows<T, YMODE, false, FULL>(A.in, loc.ia, loc.ib, A.in_es, nullptr, loc, -2 * t);
#pragma unroll
    for (int m = 0; m < E / R0; ++m)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int cj = q * L0 + TPL * m;
        C x;
        if (!trig) x = in_up.load(cj);
        else if (cj < N / 2) x = in_up.load(2 * cj);
        else {
          x = in_dn.load(2 * N - 1 - 2 * cj);
          if (neg_odd) x = {-x.x, -x.y};   // odd input rows
        }
        v[m * R0 + q] = x;
      }
  }
  const RowsOut out_up = r2_rows<T, YMODE, SPLIT, FULL>(A.out, loc.oa, loc.ob, A.out_es, A.row_tab, loc, t);
  const RowsOut out_dn = r2_rows<T, YMODE, SPLIT, FULL>(A.out, loc.oa, loc.ob, A.out_es, A.row_tab, loc, TPL - t);
  // copy the untransformed tail when the result goes to another array (never with the fused source: line_len == N there)
  if (!FUSED && A.line_len > N && (SPLIT || A.in != A.out)) {
    const RowsIn in_t = r2_rows<T, YMODE, false, FULL>(A.in, loc.ia, loc.ib, A.in_es, nullptr, loc, t);
    for (int i = N; i + t < A.line_len; i += TPL) {
      const C x = in_t.load(i);
      out_up.store(i, x.x, x.y);
    }
  }
  r2_dif_stage<T, Cfg, YMODE, 0>(v, sm, A, c, t);
  r2_group_sync<Cfg, YMODE>(c);
  if (NS > 1) { r2_dif_stage<T, Cfg, YMODE, (NS > 1 ? 1 : 0)>(v, sm, A, c, t); r2_group_sync<Cfg, YMODE>(c); }
  if (NS > 2) { r2_dif_stage<T, Cfg, YMODE, (NS > 2 ? 2 : 0)>(v, sm, A, c, t); r2_group_sync<Cfg, YMODE>(c); }
  if (NS > 3) { r2_dif_stage<T, Cfg, YMODE, (NS > 3 ? 3 : 0)>(v, sm, A, c, t); r2_group_sync<Cfg, YMODE>(c); }

  // separation pass over pairs (k, N - k): k = t + m TPL covers 0 .. N/2 - 1, thread 0 adds k = N/2
  constexpr bool rev_out = KIND == K_RODFT10;
  static_assert(N % 2 == 0 && (N / 2) % TPL == 0, "pair pass needs TPL | N/2");
  const R2Pair<T, Cfg, YMODE> P(t, c);
  const C* mak_t = A.mak + t;
#pragma unroll
  for (int m = 0; m < (N / 2) / TPL; ++m) {
    const int kc = m * TPL;                 // k = t + kc, N - k = (TPL - t) + kmc
    const int kmc = N - TPL - kc;
    const bool kpos = kc > 0 || t > 0;      // k > 0
    const C zk = sm[P.slot_k(kc)];
    const C zm = sm[P.slot_m(kc)];
    const T sx = zk.x + zm.x, dx = zk.x - zm.x, sy = zk.y + zm.y, dy = zk.y - zm.y;
    if (!trig) {
      // A_k = (sx, dy) / 2, B_k = (sy, -dx) / 2 ; halfcomplex: re at k, im at N - k
      out_up.store(kc, T(0.5) * sx, T(0.5) * sy);
      if (kpos) {
        if (!YMODE && (A.flags & CB_R2_XSPLIT)) out_up.store(N / 2 + kc, T(0.5) * dy, T(-0.5) * dx);   // N/2 + k
        else out_dn.store(kmc, T(0.5) * dy, T(-0.5) * dx);                                               // N - k
      }
    } else {
      const C cs = cx_ldg(mak_t + kc);
      const T xa = cs.x * sx + cs.y * dy, xb = cs.x * sy - cs.y * dx;
      if (rev_out) out_dn.store(kmc - 1, xa, xb);   // N - 1 - k
      else out_up.store(kc, xa, xb);
      if (kpos) {
        const T ya = cs.y * sx - cs.x * dy, yb = cs.y * sy + cs.x * dx;
        if (rev_out) out_up.store(kc - 1, ya, yb);  // k - 1
        else out_dn.store(kmc, ya, yb);             // N - k
      }
    }
  }
  if (t == 0) {
    // k = N/2 pairs with itself
    const C zk = sm[Lay::at(c, Cfg::rev(N / 2))];
    const T sx = zk.x + zk.x, sy = zk.y + zk.y;
    if (!trig) {
      out_up.store(N / 2, T(0.5) * sx, T(0.5) * sy);
    } else {
      const C cs = cx_ldg(A.mak + N / 2);
      const T xa = cs.x * sx, xb = cs.x * sy;
      out_up.store(rev_out ? N - 1 - N / 2 : N / 2, xa, xb);
      const T ya = cs.y * sx, yb = cs.y * sy;
      out_up.store(rev_out ? N / 2 - 1 : N - N / 2, ya, yb);
    }
  }
}

// ---- backward kinds: HC2R, REDFT01, RODFT01 --------------------------------------------------------
template <class T, class Cfg, bool YMODE, bool SPLIT, int KIND, bool FULL>
__global__ void __launch_bounds__(Cfg::TPL* Cfg::G, Cfg::MINB) r2r2_bwd_kernel(const R2Args<T> A) {
  static_assert(KIND == K_HC2R || KIND == K_REDFT01 || KIND == K_RODFT01, "backward kinds");
  static_assert(YMODE || r2_group_local<Cfg>(), "x mode relies on per-transform barriers");
  static_assert(!SPLIT || YMODE, "split rows exist only for the strided (y) transforms");
  using C = Cx<T>;
  using Lay = R2Lay<T, Cfg, YMODE>;
  extern __shared__ __align__(16) unsigned char cb_smem_raw[];
  C* sm = reinterpret_cast<C*>(cb_smem_raw);
  constexpr int N = Cfg::N, TPL = Cfg::TPL, E = Cfg::E, G = Cfg::G, NS = Cfg::NS;
  static_assert((TPL & (TPL - 1)) == 0, "threads per transform must be a power of two");
  const int tid = threadIdx.x;
  const int c = YMODE ? tid % G : tid / TPL;
  const int t = (YMODE ? tid / G : tid) & (TPL - 1);
  const R2Loc<T> loc = r2_locate<T, Cfg, YMODE>(A, c);
  if (!YMODE && !loc.has_a) return;
  constexpr bool trig = KIND != K_HC2R;
  constexpr bool rv = KIND == K_RODFT01;
  using RowsIn = R2Rows<T, YMODE, SPLIT, FULL>;
  using RowsOut = R2Rows<T, YMODE, false, FULL>;
  const RowsIn in_up = r2_rows<T, YMODE, SPLIT, FULL>(A.in, loc.ia, loc.ib, A.in_es, A.row_tab, loc, t);
  const RowsIn in_dn = r2_rows<T, YMODE, SPLIT, FULL>(A.in, loc.ia, loc.ib, A.in_es, A.row_tab, loc, TPL - t);

  // pre-pass: Z_k = W^a_k + i W^b_k and Z_{N-k} = conj W^a_k + i conj W^b_k, stored re/im swapped.
  // Two steps so that all global loads of a thread are in flight together.
  static_assert(N % 2 == 0 && (N / 2) % TPL == 0, "pair pass needs TPL | N/2");
  constexpr int NK = (N / 2) / TPL;
  const R2Pair<T, Cfg, YMODE> P(t, c);
  const C* mak_t = A.mak + t;
  // spectra of the two real sequences at k from the rows at k (xk) and N - k (xm), scattered to the slots of
  // Z_k and Z_{N-k}
  auto combine = [&](const C& cs, const C& xk, const C& xm, C& zk, C& zm) {
    C wa, wb;
    if (!trig) {
      wa = C{xk.x, xm.x};
      wb = C{xk.y, xm.y};
    } else {
      wa = C{cs.x * xk.x + cs.y * xm.x, cs.y * xk.x - cs.x * xm.x};
      wb = C{cs.x * xk.y + cs.y * xm.y, cs.y * xk.y - cs.x * xm.y};
    }
    zk = C{wa.y + wb.x, wa.x - wb.y};   // Z_k = (wa.x - wb.y, wa.y + wb.x), stored swapped
    zm = C{wb.x - wa.y, wa.x + wb.y};
  };
  {
    C xk[NK], xm[NK];
#pragma unroll
    for (int m = 0; m < NK; ++m) {
      const int kc = m * TPL, kmc = N - TPL - kc;
      const bool kpos = kc > 0 || t > 0;
      xm[m] = C{T(0), T(0)};
      if (!trig) {
        xk[m] = in_up.load(kc);
        if (kpos) xm[m] = (!YMODE && (A.flags & CB_R2_XSPLIT)) ? in_up.load(N / 2 + kc) : in_dn.load(kmc);
      } else {
        xk[m] = rv ? in_dn.load(kmc - 1) : in_up.load(kc);               // N - 1 - k : k
        if (kpos) xm[m] = rv ? in_up.load(kc - 1) : in_dn.load(kmc);     // k - 1 : N - k   (X_N := 0)
      }
    }
#pragma unroll
    for (int m = 0; m < NK; ++m) {
      const int kc = m * TPL;
      const bool kpos = kc > 0 || t > 0;
      C cs = C{T(1), T(0)};
      if (trig) cs = cx_ldg(mak_t + kc);
      C zk, zm;
      combine(cs, xk[m], xm[m], zk, zm);
      sm[P.slot_k(kc)] = zk;
      if (kpos) sm[P.slot_m(kc)] = zm;
    }
    if (t == 0) {
      // k = N/2: its own mirror
      C a0, a1 = C{T(0), T(0)};
      if (!trig) a0 = in_up.load(N / 2);
      else {
        a0 = in_up.load(rv ? N - 1 - N / 2 : N / 2);
        a1 = in_up.load(rv ? N / 2 - 1 : N - N / 2);
      }
      C cs = C{T(1), T(0)};
      if (trig) cs = cx_ldg(A.mak + N / 2);
      C zk, zm;
      combine(cs, a0, a1, zk, zm);
      sm[Lay::at(c, Cfg::rev(N / 2))] = zk;
    }
  }
  const RowsOut out_up = r2_rows<T, YMODE, false, FULL>(A.out, loc.oa, loc.ob, A.out_es, nullptr, loc, trig ? 2 * t : t);
  const RowsOut out_dn = r2_rows<T, YMODE, false, FULL>(A.out, loc.oa, loc.ob, A.out_es, nullptr, loc, -2 * t);
  if (A.line_len > N && (SPLIT || A.in != A.out)) {
    const RowsOut out_t = r2_rows<T, YMODE, false, FULL>(A.out, loc.oa, loc.ob, A.out_es, nullptr, loc, t);
    for (int i = N; i + t < A.line_len; i += TPL) {
      const C x = in_up.load(i);
      out_t.store(i, x.x, x.y);
    }
  }
  r2_group_sync<Cfg, YMODE>(c);
  C v[E];
  if (NS > 3) { r2_dit_stage<T, Cfg, YMODE, (NS > 3 ? 3 : 0)>(v, sm, A, c, t); r2_group_sync<Cfg, YMODE>(c); }
  if (NS > 2) { r2_dit_stage<T, Cfg, YMODE, (NS > 2 ? 2 : 0)>(v, sm, A, c, t); r2_group_sync<Cfg, YMODE>(c); }
  if (NS > 1) { r2_dit_stage<T, Cfg, YMODE, (NS > 1 ? 1 : 0)>(v, sm, A, c, t); r2_group_sync<Cfg, YMODE>(c); }
  r2_dit_stage<T, Cfg, YMODE, 0>(v, sm, A, c, t);
  // natural-order samples (swapped back: a = im, b = re) straight to global memory
  {
    constexpr int R0 = Cfg::R(0), L0 = N / R0;
#pragma unroll
    for (int m = 0; m < E / R0; ++m)
#pragma unroll
      for (int q = 0; q < R0; ++q) {
        const int cj = q * L0 + TPL * m;
        const C x = v[m * R0 + q];
        if (!trig) out_up.store(cj, x.y, x.x);
        else if (cj < N / 2) out_up.store(2 * cj, x.y, x.x);
        else if (rv) out_dn.store(2 * N - 1 - 2 * cj, -x.y, -x.x);   // odd output rows change sign
        else out_dn.store(2 * N - 1 - 2 * cj, x.y, x.x);
      }
  }
}

}  // namespace cb
