// CPU emulation of the reference-order tridiagonal kernels of cans_b200/csrc/thomas_kernels.cuh: `thomas_factor_kernel`
// (pivots in the reference's operation order, incl. the singular-pivot pin and the periodic closure's auxiliary solve,
// /root/reference/src/solver.f90:138-166,201-283) followed by `thomas_seq_kernel` (thomas_variant = 0: the two sweeps in the
// reference's order).  The kernels' own source, compiled by g++, one emulated thread per column, driven the way
// capi.cu's gaussel_prepare / gaussel_apply launch them (a cache miss on slot 0; full or x / y deduplicated pivot cache).
// The pipelined kernel (shared memory, barriers, TMA) is covered on the GPU, where it is held to this variant's result.
//
// usage: emu_thomas <f64|f32> nx ny nz n_rows periodic nopin dedup_x dedup_y norm dir
//        reads p.bin [nz][ny][nx], lam.bin [ny][nx], a.bin b.bin c.bin [>= n_rows]; writes p_out.bin
//        emu_thomas dtdma <f64|f32> nx ny nz n_rows periodic has_lam norm dir nsplit start_0 .. start_nsplit
//        the same files through the four kernels of dtdma_kernels.cuh (gaussel_dtdma, src/solver.f90:309-517)
// TEST INFRASTRUCTURE.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
// the exactly rounded intrinsics: plain IEEE operations (this file is compiled without FMA contraction)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline unsigned __float_as_uint(float d) { unsigned r; memcpy(&r, &d, 4); return r; }
// declarations the rest of the header needs to parse (kernels that are not run here)
template <class T> static T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> static T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> static T __shfl_down_sync(unsigned, T v, int) { return v; }
template <class T> static T __shfl_sync(unsigned, T v, int) { return v; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
static inline void __trap() {}
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { const unsigned long long o = *a; *a += v; return o; }
static inline int atomicAdd(int* a, int v) { const int o = *a; *a += v; return o; }
static inline int atomicOr(int* a, int v) { const int o = *a; *a |= v; return o; }
static inline int atomicExch(int* a, int v) { const int o = *a; *a = v; return o; }
template <class T> static T __ldg(const T* p) { return *p; }
#define __launch_bounds__(...)

#include "../../cans_b200/csrc/thomas_kernels.cuh"
#include "../../cans_b200/csrc/dtdma_kernels.cuh"

using namespace cb;

template <class F> static void launch(unsigned grid, unsigned block, F kernel) {
  gridDim = dim3(grid); blockDim = dim3(block);
  for (unsigned bx = 0; bx < grid; ++bx)
    for (unsigned tx = 0; tx < block; ++tx) {
      blockIdx = uint3{bx, 0, 0};
      threadIdx = uint3{tx, 0, 0};
      kernel();
    }
}
template <class T> static std::vector<T> rd(const std::string& dir, const char* name, size_t n) {
  std::vector<T> v(n);
  FILE* f = fopen((dir + "/" + name + ".bin").c_str(), "rb");
  if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "emu_thomas: cannot read %s\n", name); exit(2); }
  fclose(f);
  return v;
}

template <class T> static int run(int nx, int ny, int nz, int n_rows, int periodic, int nopin, int dedx, int dedy, double norm,
                                  const std::string& dir) {
  auto p = rd<T>(dir, "p", (size_t)nx * ny * nz);
  auto lam = rd<T>(dir, "lam", (size_t)nx * ny);
  auto a = rd<T>(dir, "a", n_rows), b = rd<T>(dir, "b", n_rows), c = rd<T>(dir, "c", n_rows);
  // make_thomas of capi.cu, natural field layout p[k][j][i]
  ThomasDev<T> D;
  D.nx = nx; D.ny = ny; D.n = n_rows; D.periodic = periodic; D.nn = periodic ? n_rows - 1 : n_rows;
  D.sj = nx; D.sk = (long long)nx * ny; D.a = a.data(); D.b = b.data(); D.c = c.data(); D.lam = lam.data(); D.lam_sj = nx;
  D.m = 1; D.chunk_layout = 2; D.xb = 0; D.xn = nx; D.out_rows = nullptr; D.nopin = nopin;
  // plan_create's deduplication extents: x keeps nx/2 + one 16-column tile in split order, y keeps rows j <= ny/2
  D.dx = dedx; D.dy = dedy;
  D.nxu = dedx ? nx / 2 + 16 : nx; D.nyu = dedy ? ny / 2 + 1 : ny;
  D.zsj = (long long)D.nn * D.nxu; D.zsk = D.nxu;
  D.dt_mode = 0; D.dt_z1 = nullptr; D.dt_rp = nullptr; D.dt_slot_small = 0; D.jb = 1;
  if (dedx && (nx % 32 != 0 || D.nxu > nx)) { fprintf(stderr, "emu_thomas: x deduplication needs whole tile pairs\n"); return 2; }
  const long long slot_z = (long long)D.nxu * D.nyu * D.nn, slot_den = (long long)D.nxu * D.nyu;
  std::vector<T> z(slot_z, (T)NAN), p2(slot_z, (T)NAN), den(slot_den, (T)NAN);
  CacheState cs;
  memset(&cs, 0, sizeof(cs));
  cs.hit = 0; cs.sel = 0; cs.nslots = 1;
  const long long ncol_s = (long long)(D.dx ? D.nxu : D.nx) * (D.dy ? D.ny / 2 + 1 : D.ny);
  launch((unsigned)((ncol_s + 127) / 128), 128, [&] { thomas_factor_kernel<T>(D, &cs, z.data(), p2.data(), den.data(), slot_z, slot_den); });
  const long long ncol = (long long)nx * ny;
  launch((unsigned)((ncol + 127) / 128), 128,
         [&] { thomas_seq_kernel<T>(D, &cs, z.data(), p2.data(), den.data(), slot_z, slot_den, p.data(), (T)norm); });
  FILE* f = fopen((dir + "/p_out.bin").c_str(), "wb");
  if (!f || fwrite(p.data(), sizeof(T), p.size(), f) != p.size()) return 2;
  fclose(f);
  return 0;
}

// gaussel_dtdma (src/solver.f90:309-517) with the z slabs of `nsplit` ranks in one address space: the four kernels of
// dtdma_kernels.cuh driven as capi.cu's gaussel_dtdma_impl drives them (cansb200_gaussel_dtdma, no coefficient cache)
template <class T> static int run_dtdma(int nx, int ny, int nz, int n_rows, int periodic, int has_lam, double norm, const std::string& dir,
                                        int nsplit, char** sv) {
  auto p = rd<T>(dir, "p", (size_t)nx * ny * nz);
  std::vector<T> lam;
  if (has_lam) lam = rd<T>(dir, "lam", (size_t)nx * ny);
  auto a = rd<T>(dir, "a", n_rows), b = rd<T>(dir, "b", n_rows), c = rd<T>(dir, "c", n_rows);
  const size_t ncol = (size_t)nx * ny;
  std::vector<T> big(3 * ncol * n_rows, (T)NAN), sm((size_t)(11 * nsplit) * ncol, (T)NAN);
  DtdmaDev<T> D;
  D.nx = nx; D.ny = ny; D.n = n_rows; D.nranks = nsplit; D.periodic = periodic;
  for (int r = 0; r <= nsplit; ++r) D.starts[r] = atoi(sv[r]);
  D.a = a.data(); D.b = b.data(); D.c = c.data(); D.lam = has_lam ? lam.data() : nullptr;
  D.st = nullptr; D.slot_big = 0; D.slot_small = 0;
  D.Z = big.data(); D.AA = big.data() + ncol * n_rows; D.CC = big.data() + 2 * ncol * n_rows;
  D.Z1 = sm.data();
  D.ra = sm.data() + ncol * nsplit; D.rc = D.ra + 2 * ncol * nsplit; D.rcw = D.rc + 2 * ncol * nsplit;
  D.rp = D.rcw + 2 * ncol * nsplit; D.rp2 = D.rp + 2 * ncol * nsplit;
  const unsigned cbk = (unsigned)((ncol + 127) / 128);
  launch(cbk, 128, [&] { dtdma_coef_kernel<T>(D); });
  launch(cbk, 128, [&] { dtdma_phase1_kernel<T>(D, p.data(), (T)norm); });
  launch(cbk, 128, [&] { dtdma_reduced_kernel<T>(D); });
  launch(7, 256, [&] { dtdma_phase3_kernel<T>(D, p.data()); });
  FILE* f = fopen((dir + "/p_out.bin").c_str(), "wb");
  if (!f || fwrite(p.data(), sizeof(T), p.size(), f) != p.size()) return 2;
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 13 && std::string(argv[1]) == "dtdma") {
    // emu_thomas dtdma <f64|f32> nx ny nz n_rows periodic has_lam norm dir nsplit start_0 .. start_nsplit
    const int nsplit = atoi(argv[11]);
    if (argc < 13 + nsplit) { fprintf(stderr, "emu_thomas dtdma: missing split starts\n"); return 2; }
    const int v[6] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8])};
    return std::string(argv[2]) == "f32" ? run_dtdma<float>(v[0], v[1], v[2], v[3], v[4], v[5], atof(argv[9]), argv[10], nsplit, argv + 12)
                                         : run_dtdma<double>(v[0], v[1], v[2], v[3], v[4], v[5], atof(argv[9]), argv[10], nsplit, argv + 12);
  }
  if (argc < 12) { fprintf(stderr, "usage: emu_thomas <f64|f32> nx ny nz n_rows periodic nopin dedup_x dedup_y norm dir\n"); return 2; }
  const int v[8] = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9])};
  const double norm = atof(argv[10]);
  return std::string(argv[1]) == "f32" ? run<float>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], norm, argv[11])
                                       : run<double>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], norm, argv[11]);
}
