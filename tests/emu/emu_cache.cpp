// CPU emulation (host-thread model of emu_threads.hpp: real barriers, warp shuffles, atomics) of the kernels that decide
// WHICH factorisation a solve uses -- `thomas_hash_kernel` (two independent content hashes of (a, b, c, lambdaxy), mirror-
// symmetry check of lambdaxy for the deduplicated cache) and `thomas_select_kernel` (slot look-up, LRU eviction) of
// cans_b200/csrc/thomas_kernels.cuh -- and of `chkdiv_kernel` (aux_kernels.cuh: the acceptance metric, warp-shuffle
// reduction + atomics), and of `r2r_direct_kernel` (fft_kernels.cuh: the O(n^2) transform of lengths / kinds no fast kernel
// serves).  The kernels' own source under g++, launched as capi.cu launches them.
//
// usage: emu_cache cache nx ny n nslots nsets dir flags...   set_<s>_{a,b,c,lam}.bin, one line "hit sel nfactor sym_bad key key2"
//                                                            per entry of the sequence given as the trailing arguments
//                                                            (each: "<set>:<nopin><dx><dy>", e.g. 0:000)
//        emu_cache chkdiv <f64|f32> n1 n2 n3 dxi dyi dir     u v w dzfi -> prints "sum max"
//        emu_cache direct <f64|f32> kind n axis nx ny nz dir arr.bin -> out.bin
// TEST INFRASTRUCTURE.
#include "emu_threads.hpp"
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
template <class T> static T __shfl_up_sync(unsigned, T v, int) { return v; }     // only in kernels that are not run here
template <class T> static T __shfl_down_sync(unsigned, T v, int) { return v; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline void __trap() {}
#include <cuda.h>
#include "../../cans_b200/csrc/aux_kernels.cuh"
#include "../../cans_b200/csrc/thomas_kernels.cuh"
#include "../../cans_b200/csrc/fft_kernels.cuh"
#include "../../cans_b200/csrc/fft_plan.hpp"

using namespace cb;

static int run_cache(int argc, char** argv) {
  const int nx = atoi(argv[2]), ny = atoi(argv[3]), n = atoi(argv[4]), nslots = atoi(argv[5]), nsets = atoi(argv[6]);
  const std::string dir = argv[7];
  std::vector<std::vector<double>> A(nsets), B(nsets), C(nsets), L(nsets);
  for (int s = 0; s < nsets; ++s) {
    const std::string p = "set_" + std::to_string(s) + "_";
    A[s] = rd<double>(dir, (p + "a").c_str(), n); B[s] = rd<double>(dir, (p + "b").c_str(), n);
    C[s] = rd<double>(dir, (p + "c").c_str(), n); L[s] = rd<double>(dir, (p + "lam").c_str(), (size_t)nx * ny);
  }
  CacheState cs;
  memset(&cs, 0, sizeof(cs));
  cs.nslots = nslots;                       // cansb200_plan_create: a zeroed state with the slot count
  for (int q = 8; q < argc; ++q) {
    const int s = atoi(argv[q]);
    const char* fl = strchr(argv[q], ':') + 1;
    ThomasDev<double> D;
    memset(&D, 0, sizeof(D));
    D.nx = nx; D.ny = ny; D.n = n; D.nn = n; D.a = A[s].data(); D.b = B[s].data(); D.c = C[s].data(); D.lam = L[s].data(); D.lam_sj = nx;
    D.nopin = fl[0] - '0'; D.dx = fl[1] - '0'; D.dy = fl[2] - '0';
    D.nxu = D.dx ? nx / 2 + 16 : nx; D.nyu = D.dy ? ny / 2 + 1 : ny;
    // gaussel_prepare (capi.cu): hash, then select
    const long long total = 3LL * D.n + (long long)nx * ny;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 592) blocks = 592;
    CacheState* st = &cs;
    launch((unsigned)blocks, 256, thomas_hash_kernel<double>, D, st);
    const unsigned long long key = cs.key_new, key2 = cs.key2_new;
    launch(1, 32, thomas_select_kernel, st);
    printf("%d %d %llu %d %llu %llu\n", cs.hit, cs.sel, cs.nfactor, cs.sym_bad, key, key2);
    if (cs.key_new != 0 || cs.key2_new != 0) return 4;   // select must leave the accumulators cleared for the next solve
  }
  return 0;
}

template <class T> static int run_chkdiv(int n1, int n2, int n3, double dxi, double dyi, const std::string& dir) {
  const size_t nh = (size_t)(n1 + 2) * (n2 + 2) * (n3 + 2);
  auto u = rd<T>(dir, "u", nh), v = rd<T>(dir, "v", nh), w = rd<T>(dir, "w", nh);
  auto dzfi = rd<T>(dir, "dzfi", n3 + 2);
  double res[2] = {0.0, 0.0};
  const long long tot = (long long)n1 * n2 * n3;
  unsigned blocks = (unsigned)((tot + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;       // cansb200_chkdiv
  if (blocks > 6) blocks = 6;                     // ... scaled to the emulator: still several CTAs and a grid-stride loop
  double* rp = res;
  launch(blocks, 256, chkdiv_kernel<T>, n1, n2, n3, (T)dxi, (T)dyi, (const T*)dzfi.data(), (const T*)u.data(), (const T*)v.data(),
         (const T*)w.data(), rp);
  printf("%.17g %.17g\n", res[0], res[1]);
  return 0;
}

// r2r_direct_kernel (fft_kernels.cuh): the O(n^2) evaluation every transform falls back to when neither the two-for-one
// kernels nor the generic engine serve its length / kind (prime n, REDFT00 / 11, RODFT00 / 11 of awkward lengths); set up as
// run_r2r (capi.cu) sets it up.  arr.bin (nz, ny, nx) -> out.bin
template <class T> static int run_direct(int kind, int n, int axis, int nx, int ny, int nz, const std::string& dir) {
  auto arr = rd<T>(dir, "arr", (size_t)nx * ny * nz);
  const int Q = slow_Q(n, kind);
  std::vector<C2<T>> cs(2 * (size_t)Q);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (long long m = 0; m < 2LL * Q; ++m) cs[m] = {(T)cosl(pi * m / Q), (T)sinl(pi * m / Q)};
  DirectArgs<T> D;
  D.n = n; D.kind = kind; D.Q = Q; D.cs = cs.data(); D.in = arr.data(); D.out = arr.data();
  if (axis == 0) { D.in_es = D.out_es = 1; D.in_ls = D.out_ls = nx; D.lines_per_group = ny; D.line_len = nx; }
  else { D.in_es = D.out_es = nx; D.in_ls = D.out_ls = 1; D.lines_per_group = nx; D.line_len = ny; }
  D.in_gs = D.out_gs = (long long)nx * ny; D.ngroups = nz;
  launch((unsigned)((long long)D.lines_per_group * nz), 128, r2r_direct_kernel<T>, D);
  wr(dir, "out", arr);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 10 && std::string(argv[1]) == "direct") {   // emu_cache direct <f64|f32> kind n axis nx ny nz dir
    const bool f32 = std::string(argv[2]) == "f32";
    const int v[6] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8])};
    return f32 ? run_direct<float>(v[0], v[1], v[2], v[3], v[4], v[5], argv[9]) : run_direct<double>(v[0], v[1], v[2], v[3], v[4], v[5], argv[9]);
  }
  if (argc >= 9 && std::string(argv[1]) == "cache") return run_cache(argc, argv);
  if (argc >= 9 && std::string(argv[1]) == "chkdiv") {
    const bool f32 = std::string(argv[2]) == "f32";
    const int n1 = atoi(argv[3]), n2 = atoi(argv[4]), n3 = atoi(argv[5]);
    return f32 ? run_chkdiv<float>(n1, n2, n3, atof(argv[6]), atof(argv[7]), argv[8]) : run_chkdiv<double>(n1, n2, n3, atof(argv[6]), atof(argv[7]), argv[8]);
  }
  fprintf(stderr, "usage: emu_cache cache nx ny n nslots nsets dir <set>:<nopin><dx><dy>... | chkdiv <f64|f32> n1 n2 n3 dxi dyi dir\n");
  return 2;
}
