// Instantiation table + launcher of the two-for-one transforms (r2r2.cuh).
// One translation unit per (precision, mode, split) includes this header and explicitly
// instantiates r2r2_run / r2r2_query, so the four units compile in parallel.
#pragma once
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
  X(1024, 0, 64, 2, 5, 16, 8, 8, 1)   \
  X(1024, 1, 128, 2, 4, 8, 8, 4, 4)   \
  X(1024, 2, 64, 4, 3, 16, 8, 8, 1)   \
  X(2048, 0, 128, 2, 1, 16, 16, 8, 1) \
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
  X(1024, 0, 64, 8, 1, 16, 8, 8, 1)   \
  X(1024, 1, 128, 4, 2, 8, 8, 4, 4)   \
  X(2048, 0, 128, 4, 1, 16, 16, 8, 1) \
  X(384, 0, 32, 8, 1, 12, 4, 4, 2)    \
  X(768, 0, 64, 4, 3, 12, 4, 4, 4)

template <class T, class Cfg, bool YMODE, bool SPLIT>
static int r2r2_launch(const R2Args<T>& A, bool fwd, cudaStream_t st) {
  using Lay = R2Lay<T, Cfg, YMODE>;
  const size_t smem = Lay::smem_bytes();
  long long grid;
  if (YMODE) grid = (long long)A.ngroups * ((A.lines_per_group + 2 * Cfg::G - 1) / (2 * Cfg::G));
  else {
    const long long npairs = ((long long)A.lines_per_group * A.ngroups + 1) / 2;
    grid = (npairs + Cfg::G - 1) / Cfg::G;
  }
  if (grid < 1) return 0;
  if (grid > 0x7fffffffLL) return -2;
  static bool attr_fwd = false, attr_bwd = false;
  if (fwd) {
    auto k = r2r2_fwd_kernel<T, Cfg, YMODE, SPLIT>;
    if (!attr_fwd) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      attr_fwd = true;
    }
    k<<<(unsigned)grid, Cfg::TPL * Cfg::G, smem, st>>>(A);
  } else {
    auto k = r2r2_bwd_kernel<T, Cfg, YMODE, SPLIT>;
    if (!attr_bwd) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      attr_bwd = true;
    }
    k<<<(unsigned)grid, Cfg::TPL * Cfg::G, smem, st>>>(A);
  }
  return 0;
}

// 0 = launched, 1 = no instantiation for this length, < 0 = error
template <class T, bool YMODE, bool SPLIT> int r2r2_run(const R2Args<T>& A, int n, int var, bool fwd, cudaStream_t st) {
#define CB_R2_CASE(N_, V_, TPL_, G_, MB_, R0_, R1_, R2_, R3_) \
  case N_ * 4 + V_: return r2r2_launch<T, R2Cfg<N_, TPL_, G_, MB_, R0_, R1_, R2_, R3_>, YMODE, SPLIT>(A, fwd, st);
  if constexpr (YMODE) {
    switch (n * 4 + var) { CB_R2_Y_CONFIGS(CB_R2_CASE) default: return 1; }
  } else {
    switch (n * 4 + var) { CB_R2_X_CONFIGS(CB_R2_CASE) default: return 1; }
  }
#undef CB_R2_CASE
}

// radices of variant `var` for length n (host-side table construction); returns the stage count or 0
template <bool YMODE> int r2r2_query(int n, int var, int radix[4]) {
#define CB_R2_CASE(N_, V_, TPL_, G_, MB_, R0_, R1_, R2_, R3_) \
  case N_ * 4 + V_: radix[0] = R0_; radix[1] = R1_; radix[2] = R2_; radix[3] = R3_; return R2Cfg<N_, TPL_, G_, MB_, R0_, R1_, R2_, R3_>::NS;
  if constexpr (YMODE) {
    switch (n * 4 + var) { CB_R2_Y_CONFIGS(CB_R2_CASE) default: return 0; }
  } else {
    switch (n * 4 + var) { CB_R2_X_CONFIGS(CB_R2_CASE) default: return 0; }
  }
#undef CB_R2_CASE
}

}  // namespace cb
