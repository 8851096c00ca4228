// CPU emulation of the two-for-one transform kernels of cans_b200/csrc/r2r2.cuh -- the hot-path FFT kernels -- from their
// own source, compiled by g++.  Every CUDA thread of a CTA is a host thread; __syncthreads / __syncwarp / bar.sync are real
// barriers between those threads (a thread that returns early drops out of them, as on the GPU); shared memory is one
// buffer per CTA; CTAs run one after the other.  Every instantiated plan of r2r2_inst.cuh (all lengths, all tuning
// variants, x and y mode, predicated and unpredicated, FP64 and FP32), the split-order variant, the fused fillps source
// (R2ArgsFill) and the SPLIT kernels with a row table can be run and are compared with the oracle by
// tests/test_emu_r2r2.py.  Launch geometry and tables are those of r2r2_launch / get_r2_tables (capi.cu).
//
// usage: emu_r2r2 <f64|f32> <x|y> n var nx ny nz dir [xsplit] [fill dxi dyi dti idx(6) val(6)] [rows]
//   arr.bin (nz, ny, nx) in; out_<kind>.bin for the six fast kinds (0 R2HC, 1 HC2R, 5 REDFT10, 4 REDFT01, 9 RODFT10,
//   8 RODFT01).  `fill`: forward x kinds only, input = haloed u.bin v.bin w.bin + dzfi.bin, output haloless.
//   `rows`: y mode through a row table (the SPLIT kernels of the distributed solve, FP64): rows.bin = int64 [size of the far
//   buffer, ny row offsets, ny group strides]; row j of group g lives at far[offset_j + g * stride_j].
// TEST INFRASTRUCTURE.
#define CB_EMU_HOST 1
#include "emu_threads.hpp"

#include "../../cans_b200/csrc/r2r2_inst.cuh"   // r2r2.cuh + the table of instantiated plans (its launchers are not used)

using namespace cb;

// twiddle tables of one plan, as get_r2_tables (capi.cu) builds them
template <class T, class Cfg> struct Tables {
  std::vector<Cx<T>> tw[4], mak;
  Tables() {
    const long double pi = 3.14159265358979323846264338327950288L;
    int Ns = Cfg::N;
    for (int s = 0; s < Cfg::NS; ++s) {
      const int Rs = Cfg::R(s), L = Ns / Rs;
      if (L > 1) {
        tw[s].resize((size_t)(Rs - 1) * L);
        for (int r = 1; r < Rs; ++r)
          for (int o = 0; o < L; ++o) {
            const long double a = -2.0L * pi * (long double)((long long)o * r) / Ns;
            tw[s][(size_t)(r - 1) * L + o] = Cx<T>{(T)cosl(a), (T)sinl(a)};
          }
      }
      Ns = L;
    }
    mak.resize((size_t)Cfg::N / 2 + 1);
    for (int k = 0; k <= Cfg::N / 2; ++k) {
      const long double a = pi * k / (2.0L * Cfg::N);
      mak[k] = Cx<T>{(T)cosl(a), (T)sinl(a)};
    }
  }
  void attach(R2Args<T>& A) const {
    for (int s = 0; s < 4; ++s) A.tw[s] = tw[s].empty() ? nullptr : tw[s].data();
    A.mak = mak.data();
  }
};

struct Opt {
  int nx, ny, nz, xsplit = 0, fill = 0, rows = 0;
  double dxi = 0, dyi = 0, dti = 0, val[6] = {0, 0, 0, 0, 0, 0};
  int idx[6] = {0, 0, 0, 0, 0, 0};
  std::string dir;
};
static const int KINDS[6] = {K_R2HC, K_HC2R, K_REDFT10, K_REDFT01, K_RODFT10, K_RODFT01};

// the geometry run_r2r (capi.cu) passes for an in-place transform of a haloless (nx, ny, nz) array
template <class T> static R2Args<T> geom(const Opt& o, bool ymode, const T* in, T* out, int kind, int n) {
  R2Args<T> A;
  A.in = in; A.out = out;
  const long long nx = o.nx, ny = o.ny;
  if (ymode) { A.in_es = A.out_es = nx; A.in_ls = A.out_ls = 1; A.lines_per_group = o.nx; A.line_len = o.ny; }
  else { A.in_es = A.out_es = 1; A.in_ls = A.out_ls = nx; A.lines_per_group = o.ny; A.line_len = o.nx; }
  A.in_gs = A.out_gs = nx * ny;
  A.ngroups = o.nz; A.kind = kind; A.row_tab = nullptr;
  A.flags = (!ymode && o.xsplit) ? CB_R2_XSPLIT : 0;
  (void)n;
  return A;
}

// EMU_PREC = 64 / 32 builds one precision per binary (the two compile side by side)
#ifndef EMU_PREC
#define EMU_PREC 0
#endif
template <class T, class Cfg, bool YMODE> static int run_cfg(const Opt& o) {
  static const Tables<T, Cfg> tab;
  const size_t nel = (size_t)o.nx * o.ny * o.nz;
  long long grid;
  if (YMODE) grid = (long long)o.nz * ((o.nx + 2 * Cfg::G - 1) / (2 * Cfg::G));
  else grid = ((((long long)o.ny * o.nz + 1) / 2) + Cfg::G - 1) / Cfg::G;
  const unsigned g = (unsigned)grid, b = Cfg::TPL * Cfg::G;
  const bool full = !YMODE || (o.nx % (2 * Cfg::G)) == 0;
  if ((YMODE ? o.ny : o.nx) != Cfg::N) { fprintf(stderr, "emu_r2r2: the transform length must equal the line length here\n"); return 2; }
  if (sizeof(cb::cb_smem_raw) < R2Lay<T, Cfg, YMODE>::smem_bytes()) { fprintf(stderr, "emu_r2r2: shared-memory buffer too small\n"); return 2; }
  if (o.fill) {
    if constexpr (!YMODE) {
      // cansb200_solve_fillps: forward x transform whose loads evaluate fillps (+ updt_rhs_b) from the haloed u, v, w
      const size_t nh = (size_t)(o.nx + 2) * (o.ny + 2) * (o.nz + 2);
      auto u = rd<T>(o.dir, "u", nh), v = rd<T>(o.dir, "v", nh), w = rd<T>(o.dir, "w", nh);
      auto dzfi = rd<T>(o.dir, "dzfi", o.nz + 2);
      const long long px = o.nx + 2, plane = px * (o.ny + 2), o111 = plane + px + 1;
      for (int kind : {K_R2HC, K_REDFT10, K_RODFT10}) {
        std::vector<T> out(nel, (T)NAN), pdummy(nh, (T)NAN);
        R2ArgsFill<T> A;
        static_cast<R2Args<T>&>(A) = geom<T>(o, false, pdummy.data() + o111, out.data(), kind, Cfg::N);
        A.in_ls = px; A.in_gs = plane;                      // the haloed source (run_r2r's gx of solve_impl)
        tab.attach(A);
        A.F.u = u.data() + o111; A.F.v = v.data() + o111; A.F.w = w.data() + o111; A.F.dzfi = dzfi.data();
        A.F.dti = (T)o.dti; A.F.dtidxi = (T)o.dti * (T)o.dxi; A.F.dtidyi = (T)o.dti * (T)o.dyi;
        A.F.sj = px; A.F.sk = plane; A.F.k0 = 0; A.F.any_rhsb = 0;
        for (int d = 0; d < 3; ++d)
          for (int s = 0; s < 2; ++s) { A.F.idx[d][s] = o.idx[2 * d + s]; A.F.val[d][s] = (T)o.val[2 * d + s]; A.F.any_rhsb |= o.idx[2 * d + s] ? 1 : 0; }
        if (kind == K_R2HC) launch(g, b, r2r2_fwd_kernel<T, Cfg, false, false, K_R2HC, true, R2ArgsFill<T>>, A);
        else if (kind == K_REDFT10) launch(g, b, r2r2_fwd_kernel<T, Cfg, false, false, K_REDFT10, true, R2ArgsFill<T>>, A);
        else launch(g, b, r2r2_fwd_kernel<T, Cfg, false, false, K_RODFT10, true, R2ArgsFill<T>>, A);
        wr(o.dir, "out_" + std::to_string(kind), out);
      }
      return 0;
    } else {
      return 2;
    }
  }
  const auto arr = rd<T>(o.dir, "arr", nel);
  std::vector<long long> rows;
  if (o.rows) rows = rd<long long>(o.dir, "rows", (size_t)2 * o.ny + 1);
  for (int kind : KINDS) {
    std::vector<T> a = arr;
    R2Args<T> A = geom<T>(o, YMODE, a.data(), a.data(), kind, Cfg::N);
    tab.attach(A);
#define EMU_CASE(KERNEL, KIND_)                                                                                  \
  case KIND_:                                                                                                    \
    if (YMODE && !full) launch(g, b, KERNEL<T, Cfg, YMODE, false, KIND_, false>, A);                             \
    else launch(g, b, KERNEL<T, Cfg, YMODE, false, KIND_, true>, A);                                             \
    break;
#define EMU_CASE_SPLIT(KERNEL, KIND_)                                                                            \
  case KIND_:                                                                                                    \
    if (!full) launch(g, b, KERNEL<T, Cfg, true, true, KIND_, false>, A);                                        \
    else launch(g, b, KERNEL<T, Cfg, true, true, KIND_, true>, A);                                               \
    break;
    if (o.rows) {
      if constexpr (YMODE && sizeof(T) == 8) {
        // SPLIT kernels: forward kinds store their result rows through the table, backward kinds load their input rows
        // through it; the far side is one buffer of rows[0] elements, row j of group g at rows[1 + j] + g * rows[1 + ny + j]
        std::vector<T> far((size_t)rows[0], (T)NAN);
        std::vector<R2Row<T>> tabr(o.ny);
        for (int j = 0; j < o.ny; ++j) { tabr[j].ptr = far.data() + rows[1 + j]; tabr[j].gs = rows[1 + o.ny + j]; }
        A.row_tab = tabr.data();
        const bool fwd = kind == K_R2HC || kind == K_REDFT10 || kind == K_RODFT10;
        if (!fwd)   // place the input rows on the far side
          for (int gq = 0; gq < o.nz; ++gq)
            for (int j = 0; j < o.ny; ++j)
              for (int i = 0; i < o.nx; ++i) tabr[j].ptr[(long long)gq * tabr[j].gs + i] = arr[((size_t)gq * o.ny + j) * o.nx + i];
        std::vector<T> local(nel, (T)NAN);
        if (fwd) { A.in = arr.data(); A.out = local.data(); }        // out is unused by a forward SPLIT kernel
        else { A.in = local.data(); A.out = a.data(); }
        switch (kind) {
          EMU_CASE_SPLIT(r2r2_fwd_kernel, K_R2HC) EMU_CASE_SPLIT(r2r2_fwd_kernel, K_REDFT10) EMU_CASE_SPLIT(r2r2_fwd_kernel, K_RODFT10)
          EMU_CASE_SPLIT(r2r2_bwd_kernel, K_HC2R) EMU_CASE_SPLIT(r2r2_bwd_kernel, K_REDFT01) EMU_CASE_SPLIT(r2r2_bwd_kernel, K_RODFT01)
        }
        if (fwd)   // gather the result rows back from the far side
          for (int gq = 0; gq < o.nz; ++gq)
            for (int j = 0; j < o.ny; ++j)
              for (int i = 0; i < o.nx; ++i) a[((size_t)gq * o.ny + j) * o.nx + i] = tabr[j].ptr[(long long)gq * tabr[j].gs + i];
      }
    } else {
      switch (kind) {
        EMU_CASE(r2r2_fwd_kernel, K_R2HC) EMU_CASE(r2r2_fwd_kernel, K_REDFT10) EMU_CASE(r2r2_fwd_kernel, K_RODFT10)
        EMU_CASE(r2r2_bwd_kernel, K_HC2R) EMU_CASE(r2r2_bwd_kernel, K_REDFT01) EMU_CASE(r2r2_bwd_kernel, K_RODFT01)
      }
    }
    wr(o.dir, "out_" + std::to_string(kind), a);
  }
  return 0;
}

template <class T, bool YMODE> static int run(int n, int var, const Opt& o) {
  // FP32: every y plan, variant 0 of the x plans (the FP64 binary covers every x variant)
#define CB_R2_CASE(N_, V_, TPL_, G_, MB_, R0_, R1_, R2_, R3_)                                                      \
  case N_ * 4 + V_:                                                                                                \
    if constexpr (sizeof(T) == 8 || YMODE || V_ == 0) return run_cfg<T, R2Cfg<N_, TPL_, G_, MB_, R0_, R1_, R2_, R3_>, YMODE>(o); \
    else return 3;
  if constexpr (YMODE && sizeof(T) == 4) {
    switch (n * 4 + var) { CB_R2_Y32_CONFIGS(CB_R2_CASE) default: return 3; }
  } else if constexpr (YMODE) {
    switch (n * 4 + var) { CB_R2_Y_CONFIGS(CB_R2_CASE) default: return 3; }
  } else {
    switch (n * 4 + var) { CB_R2_X_CONFIGS(CB_R2_CASE) default: return 3; }
  }
#undef CB_R2_CASE
}

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: emu_r2r2 <f64|f32> <x|y> n var nx ny nz dir [xsplit] [rows] [fill dxi dyi dti idx(6) val(6)]\n"); return 2; }
  const bool f32 = std::string(argv[1]) == "f32", ymode = std::string(argv[2]) == "y";
  const int n = atoi(argv[3]), var = atoi(argv[4]);
  Opt o;
  o.nx = atoi(argv[5]); o.ny = atoi(argv[6]); o.nz = atoi(argv[7]); o.dir = argv[8];
  for (int i = 9; i < argc; ++i) {
    const std::string s = argv[i];
    if (s == "xsplit") o.xsplit = 1;
    else if (s == "rows") o.rows = 1;
    else if (s == "fill") {
      if (i + 15 >= argc) { fprintf(stderr, "emu_r2r2: fill needs 15 parameters\n"); return 2; }
      o.fill = 1; o.dxi = atof(argv[i + 1]); o.dyi = atof(argv[i + 2]); o.dti = atof(argv[i + 3]);
      for (int q = 0; q < 6; ++q) { o.idx[q] = atoi(argv[i + 4 + q]); o.val[q] = atof(argv[i + 10 + q]); }
      i += 15;
    }
  }
#if EMU_PREC != 64
  if (f32) return ymode ? run<float, true>(n, var, o) : run<float, false>(n, var, o);
#endif
#if EMU_PREC != 32
  if (!f32) return ymode ? run<double, true>(n, var, o) : run<double, false>(n, var, o);
#endif
  fprintf(stderr, "emu_r2r2: this binary was built for the other precision\n");
  return 2;
}
