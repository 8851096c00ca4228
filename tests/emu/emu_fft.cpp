// CPU emulation of the r2r tile kernels: runs every phase of
// cans_b200/csrc/fft_phases.cuh for tid = 0..nthr-1 (the code a CUDA block
// runs, with __syncthreads() between phases) and checks the result against a
// long-double evaluation of the FFTW r2r definitions (FFTW manual 4.8.2-4.8.4).
//
// usage: emu_fft n kind ymode tile_lines nthr nlines [f32]
// prints "maxrel <err>" and exits non-zero if err exceeds the tolerance.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../cans_b200/csrc/fft_phases.cuh"
#include "../../cans_b200/csrc/fft_plan.hpp"

using namespace cb;
typedef long double LD;
static const LD PI = 3.14159265358979323846264338327950288L;

static void naive(int kind, int n, const std::vector<LD>& x, std::vector<LD>& y) {
  y.assign(n, 0);
  for (int k = 0; k < n; ++k) {
    LD acc = 0;
    switch (kind) {
      case K_R2HC:
        if (k <= n / 2) { for (int j = 0; j < n; ++j) acc += x[j] * cosl(2 * PI * ((long long)j * k % n) / n); }
        else { int kk = n - k; for (int j = 0; j < n; ++j) acc -= x[j] * sinl(2 * PI * ((long long)j * kk % n) / n); }
        break;
      case K_HC2R:
        acc = x[0];
        for (int f = 1; f < (n + 1) / 2; ++f)
          acc += 2 * (x[f] * cosl(2 * PI * ((long long)f * k % n) / n) - x[n - f] * sinl(2 * PI * ((long long)f * k % n) / n));
        if (n % 2 == 0) acc += (k % 2 ? -1 : 1) * x[n / 2];
        break;
      case K_REDFT10: for (int j = 0; j < n; ++j) acc += 2 * x[j] * cosl(PI * (j + 0.5L) * k / n); break;
      case K_REDFT01: acc = x[0]; for (int j = 1; j < n; ++j) acc += 2 * x[j] * cosl(PI * j * (k + 0.5L) / n); break;
      case K_RODFT10: for (int j = 0; j < n; ++j) acc += 2 * x[j] * sinl(PI * (j + 0.5L) * (k + 1) / n); break;
      case K_RODFT01:
        acc = (k % 2 ? -1 : 1) * x[n - 1];
        for (int j = 0; j < n - 1; ++j) acc += 2 * x[j] * sinl(PI * (j + 1) * (k + 0.5L) / n);
        break;
      default: break;
    }
    y[k] = acc;
  }
}

template <class T> struct DevTables {
  std::vector<C2<T>> tw, twp, mak;
  FftDev<T> dev;
  void build(const HostFftPlan& H) {
    tw.resize(H.M); twp.resize(H.M / 2 + 1); mak.resize(H.M + 1);
    for (int i = 0; i < H.M; ++i) tw[i] = {(T)H.tw_re[i], (T)H.tw_im[i]};
    for (int i = 0; i <= H.M / 2; ++i) twp[i] = {(T)H.twp_re[i], (T)H.twp_im[i]};
    for (int i = 0; i <= H.M; ++i) mak[i] = {(T)H.mak_re[i], (T)H.mak_im[i]};
    dev.n = H.n; dev.M = H.M; dev.kind = H.kind; dev.nstages = (int)H.radix.size();
    for (int s = 0; s < dev.nstages; ++s) dev.radix[s] = H.radix[s];
    dev.tw = tw.data(); dev.twp = twp.data(); dev.mak = mak.data(); dev.rev = H.rev.data();
  }
};

template <class T, class Lay>
static void run_tile(const FftArgs<T>& A, const Tile& tl, std::vector<T>& smem, const Lay& lay, int nthr) {
  T* s = smem.data();
  const bool fwd = kind_is_forward(A.P.kind);
  if (fwd) { for (int t = 0; t < nthr; ++t) phase_fwd_load(A, tl, s, lay, t, nthr); }
  else     { for (int t = 0; t < nthr; ++t) phase_bwd_pre(A, tl, s, lay, t, nthr); }
  for (int st = 0; st < A.P.nstages; ++st)
    for (int t = 0; t < nthr; ++t) phase_stage(A, tl, s, lay, st, t, nthr);
  if (fwd) { for (int t = 0; t < nthr; ++t) phase_fwd_post(A, tl, s, lay, t, nthr); }
  else     { for (int t = 0; t < nthr; ++t) phase_bwd_out(A, tl, s, lay, t, nthr); }
}

template <class T> static int run(int n, int kind, int ymode, int tile_lines, int nthr, int nlines, int tail) {
  HostFftPlan H = make_host_plan(n, kind);
  if (!H.fast) { printf("notfast\n"); return 3; }
  DevTables<T> D; D.build(H);
  const int line_len = n + tail;
  // layout: ymode -> field [g][i=line_len][l=nlines_per_group]; xmode -> [g][l][i]
  const int ngroups = 2, lpg = nlines;
  std::vector<T> in((size_t)ngroups * lpg * line_len), out(in.size(), (T)-777);
  srand(1234 + n + kind);
  for (auto& v : in) v = (T)(2.0 * rand() / RAND_MAX - 1.0);
  FftArgs<T> A;
  A.P = D.dev; A.in = in.data(); A.out = out.data();
  if (ymode) { A.in_es = A.out_es = lpg; A.in_ls = A.out_ls = 1; }
  else       { A.in_es = A.out_es = 1;   A.in_ls = A.out_ls = line_len; }
  A.in_gs = A.out_gs = (long long)lpg * line_len;
  A.lines_per_group = lpg; A.ngroups = ngroups; A.line_len = line_len; A.tile_lines = tile_lines; A.ymode = ymode;
  std::vector<T> smem(tile_smem_elems(A), (T)0);
  long long nt = num_tiles(A);
  for (long long t = 0; t < nt; ++t) {
    Tile tl = make_tile(A, t);
    if (ymode) { LayY lay{tile_lines}; run_tile(A, tl, smem, lay, nthr); }
    else { LayX lay{LayX::line_len(H.M)}; run_tile(A, tl, smem, lay, nthr); }
  }
  // check
  LD maxerr = 0, maxref = 0;
  std::vector<LD> x(n), y;
  for (int g = 0; g < ngroups; ++g)
    for (int l = 0; l < lpg; ++l) {
      for (int i = 0; i < n; ++i) x[i] = in[(size_t)g * A.in_gs + (size_t)l * A.in_ls + (size_t)i * A.in_es];
      naive(kind, n, x, y);
      for (int i = 0; i < line_len; ++i) {
        LD got = out[(size_t)g * A.out_gs + (size_t)l * A.out_ls + (size_t)i * A.out_es];
        LD ref = i < n ? y[i] : (LD)in[(size_t)g * A.in_gs + (size_t)l * A.in_ls + (size_t)i * A.in_es];
        LD e = fabsl(got - ref);
        if (e > maxerr) maxerr = e;
        if (fabsl(ref) > maxref) maxref = fabsl(ref);
      }
    }
  LD rel = maxerr / (maxref > 0 ? maxref : 1);
  printf("maxrel %.3Le  (n=%d kind=%d ymode=%d radices:", rel, n, kind, ymode);
  for (int r : H.radix) printf(" %d", r);
  printf(")\n");
  LD tol = sizeof(T) == 8 ? 5e-14L : 3e-5L;
  return rel < tol ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: emu_fft n kind ymode tile_lines nthr nlines [f32] [tail]\n"); return 2; }
  int n = atoi(argv[1]), kind = atoi(argv[2]), ymode = atoi(argv[3]), tl = atoi(argv[4]), nthr = atoi(argv[5]), nl = atoi(argv[6]);
  bool f32 = argc > 7 && atoi(argv[7]) != 0;
  int tail = argc > 8 ? atoi(argv[8]) : 0;
  return f32 ? run<float>(n, kind, ymode, tl, nthr, nl, tail) : run<double>(n, kind, ymode, tl, nthr, nl, tail);
}
