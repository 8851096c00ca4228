// Instantiation table + launcher of the two-for-one transforms (r2r2.cuh).
// One translation unit per (precision, mode, split) includes this header and explicitly
// instantiates r2r2_run / r2r2_query, so the four units compile in parallel.
#pragma once
#include <set>
#include "r2r2.cuh"

namespace cb {

// X(N, VAR, TPL, G, MINB, R0, R1, R2, R3); VAR = tuning variant (0 = default).
// x mode: G = transforms (line pairs) per CTA.
#define CB_R2_X_CONFIGS(X)            \
  X(64, 0, 4, 32, 1, 16, 4, 1, 1)     \
  X(128, 0, 8, 16, 1, 16, 8, 1, 1)    \
  X(256, 0, 16, 8, 1, 16, 16, 1, 1)   \
  X(512, 0, 64, 4, 4, 8, 8, 8, 1)     \
  X(512, 1, 32, 8, 2, 16, 8, 4, 1)    \
  X(512, 2, 64, 1, 12, 8, 8, 8, 1)    \
  X(512, 3, 64, 2, 8, 8, 8, 8, 1)     \
  X(1024, 0, 64, 1, 10, 16, 8, 8, 1)  \
  X(1024, 1, 64, 2, 5, 16, 8, 8, 1)   \
  X(1024, 2, 64, 4, 3, 16, 8, 8, 1)   \
  X(1024, 3, 128, 2, 4, 8, 8, 4, 4)   \
  X(2048, 0, 128, 1, 4, 16, 16, 8, 1) \
  X(2048, 1, 128, 2, 1, 16, 16, 8, 1) \
  X(2048, 2, 128, 2, 2, 16, 16, 8, 1) \
  X(2048, 3, 128, 1, 3, 16, 16, 8, 1) \
  X(384, 0, 32, 8, 1, 12, 4, 4, 2)    \
  X(768, 0, 64, 4, 1, 12, 4, 4, 4)
// y mode: G = column pairs per CTA (8 pairs of FP64 = one 128-byte row).
#define CB_R2_Y_CONFIGS(X)            \
  X(64, 0, 4, 8, 1, 16, 4, 1, 1)      \
  X(128, 0, 8, 8, 1, 16, 8, 1, 1)     \
  X(256, 0, 16, 8, 1, 16, 16, 1, 1)   \
  X(512, 0, 64, 8, 2, 8, 8, 8, 1)     \
  X(512, 1, 32, 8, 3, 16, 8, 4, 1)    \
  X(512, 2, 64, 4, 4, 8, 8, 8, 1)     \
  X(512, 3, 32, 8, 2, 16, 8, 4, 1)    \
  X(1024, 0, 64, 8, 1, 16, 8, 8, 1)   \
  X(1024, 1, 128, 4, 2, 8, 8, 4, 4)   \
  X(1024, 2, 32, 8, 1, 16, 16, 4, 1)  \
  X(1024, 3, 32, 4, 2, 16, 16, 4, 1)  \
  X(2048, 0, 128, 4, 1, 16, 16, 8, 1) \
  X(384, 0, 32, 8, 1, 12, 4, 4, 2)    \
  X(768, 0, 32, 8, 2, 12, 8, 8, 1)    \
  X(768, 1, 64, 8, 2, 12, 4, 4, 4)    \
  X(768, 2, 64, 4, 3, 12, 4, 4, 4)    \
  X(768, 3, 32, 4, 4, 12, 8, 8, 1)

// y mode, FP32: 16 pairs of floats = one 128-byte row
#define CB_R2_Y32_CONFIGS(X)          \
  X(64, 0, 4, 16, 1, 16, 4, 1, 1)     \
  X(128, 0, 8, 16, 1, 16, 8, 1, 1)    \
  X(256, 0, 16, 16, 1, 16, 16, 1, 1)  \
  X(512, 0, 32, 16, 2, 16, 8, 4, 1)   \
  X(512, 1, 64, 16, 1, 8, 8, 8, 1)    \
  X(512, 2, 64, 8, 2, 8, 8, 8, 1)     \
  X(1024, 0, 64, 16, 1, 16, 8, 8, 1)  \
  X(1024, 1, 128, 4, 2, 8, 8, 4, 4)   \
  X(2048, 0, 128, 8, 1, 16, 16, 8, 1) \
  X(384, 0, 32, 16, 1, 12, 4, 4, 2)   \
  X(768, 0, 32, 16, 2, 12, 8, 8, 1)   \
  X(768, 1, 64, 8, 2, 12, 4, 4, 4)

extern int g_r2_default_carveout;

template <class K>
static int r2r2_launch_kernel(K k, const void* args, unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool split = false) {
  // one attribute call per kernel and process (keyed by the function address)
  static std::set<const void*> done;
  if (!done.count((const void*)k)) {
    // the SPLIT kernels may be launched with padded shared memory (R2Args::smem_pad): allow the maximum once
    const size_t lim = split && smem < 200 * 1024 ? 200 * 1024 : smem;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim) != cudaSuccess) return -3;
    if (!g_r2_default_carveout)
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    done.insert((const void*)k);
  }
  void* kargs[1] = {const_cast<void*>(args)};
  return cudaLaunchKernel((const void*)k, dim3(grid), dim3(block), kargs, smem, st) == cudaSuccess ? 0 : -4;
}

template <class T, class Cfg, bool YMODE, bool SPLIT>
static int r2r2_launch(const R2Args<T>& A, bool fwd, cudaStream_t st) {
  using Lay = R2Lay<T, Cfg, YMODE>;
  size_t smem = Lay::smem_bytes();
  if (SPLIT && A.smem_pad > 0 && smem + (size_t)A.smem_pad <= 200 * 1024) smem += (size_t)A.smem_pad;
  long long grid;
  if (YMODE) grid = (long long)A.ngroups * ((A.lines_per_group + 2 * Cfg::G - 1) / (2 * Cfg::G));
  else {
    const long long npairs = ((long long)A.lines_per_group * A.ngroups + 1) / 2;
    grid = (npairs + Cfg::G - 1) / Cfg::G;
  }
  if (grid < 1) return 0;
  if (grid > 0x7fffffffLL) return -2;
  if (!YMODE && (long long)A.lines_per_group * A.ngroups > 0x7ffffff0LL) return -2;   // 32-bit line indices (r2_locate)
  const unsigned g = (unsigned)grid, b = Cfg::TPL * Cfg::G;
  // y mode: the unpredicated kernels need every column pair of every CTA to exist
  const bool full = !YMODE || (A.lines_per_group % (2 * Cfg::G)) == 0;
#define CB_R2_LAUNCH(KERNEL, KIND_)                                                                        \
  case KIND_:                                                                                              \
    if constexpr (YMODE) {                                                                                 \
      if (!full) return r2r2_launch_kernel(KERNEL<T, Cfg, YMODE, SPLIT, KIND_, false>, &A, g, b, smem, st, SPLIT); \
    }                                                                                                      \
    return r2r2_launch_kernel(KERNEL<T, Cfg, YMODE, SPLIT, KIND_, true>, &A, g, b, smem, st, SPLIT);
  switch (A.kind) {
    CB_R2_LAUNCH(r2r2_fwd_kernel, K_R2HC)
    CB_R2_LAUNCH(r2r2_fwd_kernel, K_REDFT10)
    CB_R2_LAUNCH(r2r2_fwd_kernel, K_RODFT10)
    CB_R2_LAUNCH(r2r2_bwd_kernel, K_HC2R)
    CB_R2_LAUNCH(r2r2_bwd_kernel, K_REDFT01)
    CB_R2_LAUNCH(r2r2_bwd_kernel, K_RODFT01)
    default: return -5;
  }
#undef CB_R2_LAUNCH
  (void)fwd;
}

// 0 = launched, 1 = no instantiation for this length, < 0 = error
template <class T, bool YMODE, bool SPLIT> int r2r2_run(const R2Args<T>& A, int n, int var, bool fwd, cudaStream_t st) {
#define CB_R2_CASE(N_, V_, TPL_, G_, MB_, R0_, R1_, R2_, R3_) \
  case N_ * 4 + V_: return r2r2_launch<T, R2Cfg<N_, TPL_, G_, MB_, R0_, R1_, R2_, R3_>, YMODE, SPLIT>(A, fwd, st);
  if constexpr (YMODE && sizeof(T) == 4) {
    switch (n * 4 + var) { CB_R2_Y32_CONFIGS(CB_R2_CASE) default: return 1; }
  } else if constexpr (YMODE) {
    switch (n * 4 + var) { CB_R2_Y_CONFIGS(CB_R2_CASE) default: return 1; }
  } else {
    switch (n * 4 + var) { CB_R2_X_CONFIGS(CB_R2_CASE) default: return 1; }
  }
#undef CB_R2_CASE
}

// forward x transform fed by the fused fillps source (R2Fill): variant 0 of every instantiated length.
// 0 = launched, 1 = no instantiation for this length, < 0 = error
template <class T, class Cfg> static int r2r2_launch_fill(const R2Args<T>& A0, const R2Fill<T>& F, cudaStream_t st) {
  R2ArgsFill<T> A;
  static_cast<R2Args<T>&>(A) = A0;
  A.F = F;
  using Lay = R2Lay<T, Cfg, false>;
  const size_t smem = Lay::smem_bytes();
  const long long npairs = ((long long)A.lines_per_group * A.ngroups + 1) / 2;
  const long long grid = (npairs + Cfg::G - 1) / Cfg::G;
  if (grid < 1) return 0;
  if (grid > 0x7fffffffLL || (long long)A.lines_per_group * A.ngroups > 0x7ffffff0LL) return -2;
  if (A.line_len != Cfg::N) return 1;   // the fused source has no untransformed tail
  auto launch = [&](auto k) -> int {
    static std::set<const void*> done;
    if (!done.count((const void*)k)) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
      done.insert((const void*)k);
    }
    void* kargs[1] = {&A};
    return cudaLaunchKernel((const void*)k, dim3((unsigned)grid), dim3(Cfg::TPL * Cfg::G), kargs, smem, st) == cudaSuccess ? 0 : -4;
  };
  switch (A.kind) {
    case K_R2HC: return launch(r2r2_fwd_kernel<T, Cfg, false, false, K_R2HC, true, R2ArgsFill<T>>);
    case K_REDFT10: return launch(r2r2_fwd_kernel<T, Cfg, false, false, K_REDFT10, true, R2ArgsFill<T>>);
    case K_RODFT10: return launch(r2r2_fwd_kernel<T, Cfg, false, false, K_RODFT10, true, R2ArgsFill<T>>);
    default: return -5;
  }
}
template <class T> int r2r2_run_fillps(const R2Args<T>& A, const R2Fill<T>& F, int n, cudaStream_t st) {
#define CB_R2_CASE(N_, V_, TPL_, G_, MB_, R0_, R1_, R2_, R3_) \
  case N_ * 4 + V_:                                           \
    if constexpr (V_ == 0) return r2r2_launch_fill<T, R2Cfg<N_, TPL_, G_, MB_, R0_, R1_, R2_, R3_>>(A, F, st); \
    else return 1;
  switch (n * 4) { CB_R2_X_CONFIGS(CB_R2_CASE) default: return 1; }
#undef CB_R2_CASE
}

// radices of variant `var` for length n (host-side table construction); returns the stage count or 0
template <bool YMODE, bool F32> int r2r2_query(int n, int var, int radix[4]) {
#define CB_R2_CASE(N_, V_, TPL_, G_, MB_, R0_, R1_, R2_, R3_) \
  case N_ * 4 + V_: radix[0] = R0_; radix[1] = R1_; radix[2] = R2_; radix[3] = R3_; return R2Cfg<N_, TPL_, G_, MB_, R0_, R1_, R2_, R3_>::NS;
  if constexpr (YMODE && F32) {
    switch (n * 4 + var) { CB_R2_Y32_CONFIGS(CB_R2_CASE) default: return 0; }
  } else if constexpr (YMODE) {
    switch (n * 4 + var) { CB_R2_Y_CONFIGS(CB_R2_CASE) default: return 0; }
  } else {
    switch (n * 4 + var) { CB_R2_X_CONFIGS(CB_R2_CASE) default: return 0; }
  }
#undef CB_R2_CASE
}

}  // namespace cb
