// CPU emulation of the step kernels of cans_b200/csrc/aux_kernels.cuh: the kernels' own source, compiled by g++, run for
// every (block, thread) of the launch geometry the library uses.  Only kernels without barriers / shuffles run here
// (fillps, correc, updt_rhs_b, lambda_unpack, fill_hash); chkdiv (warp shuffles + atomics) is covered on the GPU.
//
// usage: emu_aux <op> <n1> <n2> <n3> <dir> [params...]     arrays are raw little-endian files in <dir> (numpy tofile)
//   fillps3d | fillps_flat   dxi dyi dti              u v w dzfi      -> p_out
//   correc3d | correc_flat   dxi dyi dt               p u v w dzci    -> u_out v_out w_out
//   updt_rhs_b               idx(6) val(6)            p               -> p_out
//   lambda_unpack            px py sx                 lam (n1 x n2)   -> lam_out
//   fill_hash                o1 o2 o3 ng1 ng2 nh seed                 -> p_out
//   fill_source              dxi dyi dti idx(6) val(6) kc   u v w dzfi -> p_out   (R2Fill / R2FillLine of r2_fill.cuh: the loads of
//                            the fused forward x transform, evaluated for every interior point as two launches on the z
//                            chunks [0, kc) and [kc, n3), set up the way capi.cu's run_r2r sets them up)
// TEST INFRASTRUCTURE (the product launches the same kernels through capi.cu).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>   // dim3, uint3 (host declarations only: g++ never sees device code)

// the execution context of the thread being emulated
static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
// device intrinsics the header mentions in kernels that are NOT run here (chkdiv): declarations so that it parses
template <class T> static T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long c, unsigned long long v) {
  const unsigned long long o = *a;
  if (o == c) *a = v;
  return o;
}
static inline double atomicAdd(double* a, double v) { const double o = *a; *a += v; return o; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }

template <class T> static T __ldg(const T* p) { return *p; }

#include "../../cans_b200/csrc/aux_kernels.cuh"
#include "../../cans_b200/csrc/r2_fill.cuh"

using namespace cb;

template <class F> static void launch(dim3 grid, dim3 block, F kernel) {
  gridDim = grid; blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              blockIdx = uint3{bx, by, bz};
              threadIdx = uint3{tx, ty, tz};
              kernel();
            }
}

template <class T> static std::vector<T> rd(const std::string& dir, const char* name, size_t n) {
  std::vector<T> v(n);
  FILE* f = fopen((dir + "/" + name + ".bin").c_str(), "rb");
  if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "emu_aux: cannot read %s\n", name); exit(2); }
  fclose(f);
  return v;
}
template <class T> static void wr(const std::string& dir, const char* name, const std::vector<T>& v) {
  FILE* f = fopen((dir + "/" + name + ".bin").c_str(), "wb");
  if (!f || fwrite(v.data(), sizeof(T), v.size(), f) != v.size()) { fprintf(stderr, "emu_aux: cannot write %s\n", name); exit(2); }
  fclose(f);
}

template <class T> static int run(const std::string& op, int n1, int n2, int n3, const std::string& dir, char** a, int na) {
  const size_t nh = (size_t)(n1 + 2) * (n2 + 2) * (n3 + 2);
  auto num = [&](int i) { if (i >= na) { fprintf(stderr, "emu_aux: missing parameter\n"); exit(2); } return atof(a[i]); };
  dim3 g, b;
  if (op == "fillps3d" || op == "fillps_flat") {
    auto u = rd<T>(dir, "u", nh), v = rd<T>(dir, "v", nh), w = rd<T>(dir, "w", nh), p = rd<T>(dir, "p", nh);
    auto dzfi = rd<T>(dir, "dzfi", n3 + 2);
    const T dxi = (T)num(0), dyi = (T)num(1), dti = (T)num(2);
    if (op == "fillps3d") {
      if (!aux_geom(n1, n2, n3, g, b)) return 3;
      launch(g, b, [&] { fillps3d_kernel<T>(n1, n2, dxi, dyi, dzfi.data(), dti, u.data(), v.data(), w.data(), p.data()); });
    } else {
      const long long tot = (long long)n1 * n2 * n3;
      launch(dim3((unsigned)((tot + 255) / 256)), dim3(256), [&] {
        fillps_kernel<T>(n1, n2, n3, dxi, dyi, dzfi.data(), dti, u.data(), v.data(), w.data(), p.data());
      });
    }
    wr(dir, "p_out", p);
  } else if (op == "correc3d" || op == "correc_flat") {
    auto u = rd<T>(dir, "u", nh), v = rd<T>(dir, "v", nh), w = rd<T>(dir, "w", nh), p = rd<T>(dir, "p", nh);
    auto dzci = rd<T>(dir, "dzci", n3 + 2);
    const T dxi = (T)num(0), dyi = (T)num(1), dt = (T)num(2);
    if (op == "correc3d") {
      if (!aux_geom(n1 + 2, n2 + 2, n3 + 2, g, b)) return 3;
      launch(g, b, [&] { correc3d_kernel<T>(n1, n2, n3, dxi, dyi, dzci.data(), dt, p.data(), u.data(), v.data(), w.data()); });
    } else {
      launch(dim3((unsigned)((nh + 255) / 256)), dim3(256), [&] {
        correc_kernel<T>(n1, n2, n3, dxi, dyi, dzci.data(), dt, p.data(), u.data(), v.data(), w.data());
      });
    }
    wr(dir, "u_out", u); wr(dir, "v_out", v); wr(dir, "w_out", w);
  } else if (op == "updt_rhs_b") {
    auto p = rd<T>(dir, "p", nh);
    RhsbPlanes B;
    for (int d = 0; d < 3; ++d)
      for (int s = 0; s < 2; ++s) { B.idx[d][s] = (int)num(2 * d + s); B.val[d][s] = num(6 + 2 * d + s); }
    for (int d = 0; d < 3; ++d) {   // one launch per direction, as cansb200_updt_rhs_b does
      if (!B.idx[d][0] && !B.idx[d][1]) continue;
      const long long face = d == 0 ? (long long)n2 * n3 : (d == 1 ? (long long)n1 * n3 : (long long)n1 * n2);
      launch(dim3((unsigned)((face + 255) / 256)), dim3(256), [&] { updt_rhs_b_kernel<T>(p.data(), n1, n2, n3, B, d); });
    }
    wr(dir, "p_out", p);
  } else if (op == "lambda_unpack") {
    auto lam = rd<T>(dir, "lam", (size_t)n1 * n2);
    std::vector<T> out((size_t)n1 * n2);
    launch(dim3(3), dim3(64), [&] { lambda_unpack_kernel<T>(lam.data(), out.data(), n1, n2, (int)num(0), (int)num(1), (int)num(2)); });
    wr(dir, "lam_out", out);
  } else if (op == "fill_source") {
    auto u = rd<T>(dir, "u", nh), v = rd<T>(dir, "v", nh), w = rd<T>(dir, "w", nh);
    auto dzfi = rd<T>(dir, "dzfi", n3 + 2);
    std::vector<T> p(nh, (T)3.25);
    const long long px = n1 + 2, plane = px * (n2 + 2), o111 = plane + px + 1;
    const int kc = (int)num(15);
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int k0 = chunk ? kc : 0, nk = chunk ? n3 - kc : kc;
      const long long delta = (long long)k0 * plane;   // `in - fs.pin` of the launch
      R2Fill<T> F;
      F.u = u.data() + o111 + delta; F.v = v.data() + o111 + delta; F.w = w.data() + o111 + delta;
      F.dzfi = dzfi.data();
      F.dti = (T)num(2); F.dtidxi = (T)num(2) * (T)num(0); F.dtidyi = (T)num(2) * (T)num(1);
      F.sj = px; F.sk = plane; F.k0 = k0;
      F.any_rhsb = 0;
      for (int d = 0; d < 3; ++d)
        for (int sd = 0; sd < 2; ++sd) {
          F.idx[d][sd] = (int)num(3 + 2 * d + sd); F.val[d][sd] = (T)num(9 + 2 * d + sd);
          F.any_rhsb |= F.idx[d][sd] ? 1 : 0;
        }
      for (int g = 0; g < nk; ++g)
        for (int j = 0; j < n2; ++j) {
          const long long off = (long long)g * plane + (long long)j * px;   // R2Loc::ia of line (g, j)
          const R2FillLine<T> L(F, off, g, j);
          for (int i = 0; i < n1; ++i) p[o111 + delta + off + i] = L.at(F, i);
        }
    }
    wr(dir, "p_out", p);
  } else if (op == "fill_hash") {
    const int nhalo = (int)num(5);
    std::vector<T> p((size_t)(n1 + 2 * nhalo) * (n2 + 2 * nhalo) * (n3 + 2 * nhalo), (T)7);
    launch(dim3(5), dim3(128), [&] {
      fill_hash_kernel<T>(p.data(), n1, n2, n3, (int)num(0), (int)num(1), (int)num(2), (int)num(3), (int)num(4), nhalo,
                          strtoull(a[6], nullptr, 10));
    });
    wr(dir, "p_out", p);
  } else {
    fprintf(stderr, "emu_aux: unknown op %s\n", op.c_str());
    return 2;
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: emu_aux <f64|f32> <op> n1 n2 n3 dir [params]\n"); return 2; }
  const std::string prec = argv[1], op = argv[2], dir = argv[6];
  const int n1 = atoi(argv[3]), n2 = atoi(argv[4]), n3 = atoi(argv[5]);
  return prec == "f32" ? run<float>(op, n1, n2, n3, dir, argv + 7, argc - 7) : run<double>(op, n1, n2, n3, dir, argv + 7, argc - 7);
}
