// C ABI of cans_b200 (include/cans_b200.h): context, plans, the solve.
// Host-side orchestration only; all arithmetic lives in the kernels of
// fft_kernels.cuh / thomas_kernels.cuh / aux_kernels.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/cans_b200.h"
#include "aux_kernels.cuh"
#include "dist_kernels.cuh"
#include "dtdma_kernels.cuh"
#include "fft_kernels.cuh"
#include "fft_plan.hpp"
#include "r2r2.cuh"
#include "thomas_kernels.cuh"

using namespace cb;

namespace cb {
// explicit instantiations live in r2r2_{x,y}{64,32}.cu
template <class T, bool YMODE, bool SPLIT> int r2r2_run(const R2Args<T>& A, int n, int var, bool fwd, cudaStream_t st);
template <bool YMODE, bool F32> int r2r2_query(int n, int var, int radix[4]);
extern template int r2r2_run<double, false, false>(const R2Args<double>&, int, int, bool, cudaStream_t);
extern template int r2r2_run<double, true, false>(const R2Args<double>&, int, int, bool, cudaStream_t);
extern template int r2r2_run<float, false, false>(const R2Args<float>&, int, int, bool, cudaStream_t);
extern template int r2r2_run<float, true, false>(const R2Args<float>&, int, int, bool, cudaStream_t);
extern template int r2r2_run<double, true, true>(const R2Args<double>&, int, int, bool, cudaStream_t);
extern template int r2r2_run<float, true, true>(const R2Args<float>&, int, int, bool, cudaStream_t);
// forward x transform fed by the fused fillps source (r2r2_xf{64,32}.cu)
template <class T> int r2r2_run_fillps(const R2Args<T>& A, const R2Fill<T>& F, int n, cudaStream_t st);
extern template int r2r2_run_fillps<double>(const R2Args<double>&, const R2Fill<double>&, int, cudaStream_t);
extern template int r2r2_run_fillps<float>(const R2Args<float>&, const R2Fill<float>&, int, cudaStream_t);
extern int g_r2_default_carveout;   // 1: do not force the maximum shared-memory carveout (capi.cu owns it)
extern template int r2r2_query<false, false>(int, int, int[4]);
extern template int r2r2_query<true, false>(int, int, int[4]);
extern template int r2r2_query<false, true>(int, int, int[4]);
extern template int r2r2_query<true, true>(int, int, int[4]);
}  // namespace cb

namespace cb { int g_r2_default_carveout = 1; }
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
// integer names of the live plans (arrplan is `integer, dimension(2,2)` on the reference's CUDA build)
static std::mutex g_plan_mu;
static std::vector<cansb200_plan*> g_plan_ids;   // id - 1 -> plan (nullptr = free)

// NVTX ranges around the solve and its stages (the reference brackets its steps with nvtxStartRange / nvtxEndRange through
// timer_tic / timer_toc, src/timer.f90:113-216, src/nvtx.f90:48-64).  NVTX3 is header-only and costs a few ns without a tool.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(CANSB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// Host -> device uploads of tables and initial values.  cudaMemcpy from pageable memory may return before the DMA has
// reached its destination, and it is only ordered against the legacy default stream -- not against the caller's
// (possibly non-blocking) stream the consuming kernel is launched on a few microseconds later.  Waiting for the legacy
// stream closes that window (it does not wait for non-blocking streams, so device-side waits of other ranks cannot block it).
static cudaError_t upload(void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(cudaStreamLegacy);
}
static cudaError_t clear_dev(void* dst, size_t bytes) {
  cudaError_t e = cudaMemset(dst, 0, bytes);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(cudaStreamLegacy);
}

// ---------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t b) {
    if (b <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    if (cudaMalloc(&p, b) != cudaSuccess) return -1;
    bytes = b;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }   // error paths of plan / context creation must not leak device memory
};

template <class T> struct FftTables {
  HostFftPlan H;
  C2<T>*tw = nullptr, *twp = nullptr, *mak = nullptr;
  uint16_t* rev = nullptr;
};
template <class T> struct DirectTables {
  int Q = 0;
  C2<T>* cs = nullptr;
};

// tables of the two-for-one fast path (r2r2.cuh), one set per (length, mode)
template <class T> struct R2Tables {
  Cx<T>* tw[4] = {nullptr, nullptr, nullptr, nullptr};
  Cx<T>* mak = nullptr;
};

struct cansb200_ctx {
  int ng[3], dims[2], ipencil_axis, rank, nranks, is_fp32;
  int n[3], lo[3], n_z[3], lo_z[3];
  size_t esz;
  int num_sms = 148;
  DevBuf scratch;   // haloless field buffer A (x pencil)
  DevBuf scratch2;  // z-major copy B[j][k][i] of the middle stages (zmajor)
  DevBuf work2;     // third pencil-sized buffer lent to the host (cansb200_get_work)
  int dtdma_tiled = 1;   // distributed TDMA: slab-local elimination on chip (pipelined kernel) instead of per-column sweeps through HBM
  int dtdma = 0;    // several ranks: keep z decomposed in the tridiagonal stage (gaussel_dtdma) instead of transposing to z pencils
  int zmajor = 1;   // 1: the y transforms write / read B, so that every row stream of the tridiagonal stage is 8 KB-strided
  DevBuf staging;   // haloed p when the caller's p is host memory
  DevBuf coef;      // a, b, c, lambdaxy staged from the host
  DevBuf zero_lam;  // all-zero lambdaxy of the z-only solve (solver_gaussel_z)
  DevBuf lam_perm;  // lambdaxy brought from the _OPENACC (packed) order to halfcomplex order (CANSB200 option lambda_order = 1)
  DevBuf ytab_fwd_pk, ytab_bwd_pk;   // row tables of the distributed y transforms when the y rows are dealt out in packed order
  bool tabs_pk = false;
  // fftini records (cansb200_fftini): the reference's fftini sees the x / y boundary conditions only; the z variant of the
  // plan is known at the first `solver` call, so the real plan is created then, one per (record, z BCs, c_or_f(3))
  struct FftiniRec { char bcxy[4]; char cf[2]; bool live; std::map<std::string, cansb200_plan*> byz; };
  std::vector<FftiniRec> fftini_recs;
  std::map<std::string, cansb200_plan*> zplans;   // z-only plans created on demand by cansb200_solve_z_bc, keyed by bcz + c_or_f(3)
  std::map<int, FftTables<double>> tabs64;
  std::map<int, FftTables<float>> tabs32;
  std::map<long long, DirectTables<double>> dtabs64;
  std::map<long long, DirectTables<float>> dtabs32;
  std::map<int, R2Tables<double>> r2tabs64;   // key = (n * 4 + variant) * 2 + ymode
  std::map<int, R2Tables<float>> r2tabs32;
  int force_generic = 0;                       // tests: route every transform through the generic engine
  int r2_variant[2] = {-1, -1};                // tuning variant of the fast path, [x, y]; -1 = per-kind default (r2_auto_variant)
  int r2_flags = 0;                            // cache hints of the fast path (CANSB200_CTX_R2_FLAGS)
  // ---- z-slab decomposition over the GPUs of one box (dims = [1, P]); see dist_kernels.cuh
  std::vector<int> ys, zs;                     // split starts of y and z, size P + 1
  void* region = nullptr;                      // IPC-exported: [flags][C = z pencil][XB = way-back buffer]
  size_t region_bytes = 0, off_C = 0, off_XB = 0;
  std::vector<void*> peer;                     // peer-mapped region bases (own rank: region)
  std::vector<size_t> peer_off_XB;             // offset of XB inside every rank's region
  bool connected = false;
  bool local_peers = false;                    // peers are contexts of this process on this device (cansb200_dist_connect_local)
  unsigned long long seq = 0;                  // sequence number of the distributed solves = value of every flag of a solve
  int* dist_status = nullptr;                  // host-mapped word set by a device-side wait that timed out
  int* dist_status_dev = nullptr;              // its device-resident twin
  int* pinned_word = nullptr;                  // page-locked scratch word for the rare device -> host read-backs
  int dist_windows = -1;                       // x windows of the pipelined exchange (-1 = auto: up to 4)
  int dist_thomas_ctas = -1;                   // CTAs of the tridiagonal kernel while it shares the GPU with the y transforms (-1 = auto)
  int dist_mode = -1;                          // exchange flavour: 0 = stores of the producing kernels, 1 = copy engines, -1 = auto
  int dist_chunks = -1;                        // z chunks of the forward half (-1 = auto)
  int dist_split_pad = -1;                     // KB of shared-memory padding of the forward SPLIT kernels (-1 = auto: one CTA per SM)
  DevBuf sendb;                                // copy-engine exchange: way-back send buffer [dest][j][k][i] (the forward one is scratch2)
  DevBuf ytab_fwd_loc, ytab_fwd_loc_pk, ztab_loc;   // ... row tables that point into the local send buffers
  std::vector<cudaStream_t> dist_cs;           // ... one copy stream per peer
  cudaStream_t dist_sF = nullptr;              // two-half schedule: stream of the forward y transforms
  cudaStream_t dist_sT = nullptr, dist_sB = nullptr;   // pipeline stages: tridiagonal solve / backward y transform (forward = caller's stream)
  std::vector<cudaEvent_t> dist_ev;            // [w] forward window done, [W + w] tridiagonal window done, [2 W] backward done
  DevBuf ytab_fwd, ytab_bwd, ztab;
  // L2-resident chain: fft-y -> tridiagonal -> ifft-y run per window of `chain_cols` x columns, windows
  // round-robin on auxiliary streams, so that the two intermediate fields never leave the L2 cache
  int chain_cols = -1;                         // 0 = off (three full-field passes), -1 = auto (two half-width windows)
  int chain_nstreams = 2;
  std::vector<cudaStream_t> aux;
  std::vector<cudaEvent_t> aux_done;
  cudaEvent_t fork_ev = nullptr;
  unsigned long long launches = 0;
  // One solve at a time per context: scratch, staging and the exchange region are shared by its plans.  The device side
  // is made safe here: a solve enqueued on another stream than the previous one first waits for that one's end.
  cudaStream_t last_stream = nullptr;
  cudaEvent_t last_done = nullptr;
  bool have_last = false;
  int cur_xsplit = 0;                           // set for the duration of a solve whose plan keeps x in split order (CB_R2_XSPLIT)
  // cansb200_solve_fillps: for the duration of that solve the forward x transform evaluates its samples from u, v, w
  struct FuseSrc {
    const void *u, *v, *w, *dzfi;               // device pointers; u, v, w haloed like p
    const void* pin;                            // p(1,1,1) of the solve: launches on z chunks are located relative to it
    double dti, dxi, dyi;
    RhsbPlanes B;                               // wall terms of updt_rhs_b (idx = 0: none)
    int any_rhsb;
  };
  const FuseSrc* fuse = nullptr;
  int fuse_fillps = 1;                          // CANSB200_CTX_FUSE_FILLPS
  int aux_3d = 1;                               // CANSB200_CTX_AUX_3D: fillps / correc with the 3-D launch geometry
  int cta_cap = 0;                              // > 0: the persistent tridiagonal kernel uses at most this many CTAs (pipelined exchange)
  int nplans = 0;                               // live plans (some switches are only legal before the first one)
  // host-memory mode: the z planes travel in chunks on two copy streams so that the x / y transforms of a chunk
  // overlap the PCIe transfer of the others (only the tridiagonal stage needs the whole field)
  int host_chunks = 16;
  int pin_host = 0;                             // 1: page-lock the caller's host array on first use (cudaHostRegister)
  std::vector<std::pair<void*, size_t>> pinned; // what we registered (released in cansb200_finalize)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  // optional per-stage CUDA-event timing (bench.py's live roofline measurement)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;   // 7 events per profiled solve
  double prof_ms[8] = {0};
  unsigned long long prof_n = 0;
};

static void prof_mark(cansb200_ctx* c, cudaStream_t st) {
  if (!c->profiling) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  c->prof_events.push_back(e);
}

struct cansb200_plan {
  cansb200_ctx* ctx;
  char bc[6], cf[3];
  int kind[2][2];   // [dir][fwd|bwd]
  int nt[2];        // transform lengths (n - ix, n - iy)
  int q, periodic_z;
  cansb200_options opt;
  // thomas
  int th_n, th_nn, th_m, th_variant, th_mmax, th_cols, th_cl, nslots;
  long long slot_z, slot_den;
  int jb = 1, th_m_b = 0, th_mmax_b = 0;  // shallow grids: y rows per tall tile of the z-major one-GPU solve, and its chunking
  int dx = 0, dy = 0, nxu = 0, nyu = 0;   // pivot-cache deduplication (ThomasDev::dx ..): decided at plan creation, verified at the first solve
  bool sym_checked = false;
  DevBuf zcache, p2cache, dencache, state;
  DevBuf dtdma_big, dtdma_small;   // distributed-TDMA coefficients (Z, AA, CC) and reduced system (cansb200_gaussel_dtdma)
  const void* zsrc_base = nullptr;  // pivot source of the pipelined kernel when it is not the plan's cache (distributed TDMA: the Z array)
  long long zsrc_slot = 0;
  DevBuf dtdma_rows;               // ... gathered coefficient rows of all ranks, per cache slot (solve_dist_dtdma)
  unsigned long long solves = 0;
  int last_fused = 0;              // did the last cansb200_solve_fillps on this plan run the fused forward x transform?
  // TMA descriptors of the pipelined substitution: pivots (per plan) and right-hand sides (per field pointer / shape)
  bool use_tma = true;
  CUtensorMap map_z;
  unsigned long long map_z_key[4] = {0, 0, 0, 0};
  struct PMap { const void* p; long long sj, sk; int nx, ny, nn, box_rows; CUtensorMap m; };
  std::vector<PMap> map_p;
};

template <class T> static std::map<int, FftTables<T>>& tabmap(cansb200_ctx* c);
template <> std::map<int, FftTables<double>>& tabmap<double>(cansb200_ctx* c) { return c->tabs64; }
template <> std::map<int, FftTables<float>>& tabmap<float>(cansb200_ctx* c) { return c->tabs32; }
template <class T> static std::map<long long, DirectTables<T>>& dtabmap(cansb200_ctx* c);
template <> std::map<long long, DirectTables<double>>& dtabmap<double>(cansb200_ctx* c) { return c->dtabs64; }
template <> std::map<long long, DirectTables<float>>& dtabmap<float>(cansb200_ctx* c) { return c->dtabs32; }

template <class T> static std::map<int, R2Tables<T>>& r2tabmap(cansb200_ctx* c);
template <> std::map<int, R2Tables<double>>& r2tabmap<double>(cansb200_ctx* c) { return c->r2tabs64; }
template <> std::map<int, R2Tables<float>>& r2tabmap<float>(cansb200_ctx* c) { return c->r2tabs32; }

// twiddles of the fast path: stage s holds w_{Ns}^{o r} at [(r-1) L + o]; mak[k] = (cos, sin)(pi k / 2n)
template <class T> static int get_r2_tables(cansb200_ctx* ctx, int n, int ymode, int& var, R2Tables<T>** out) {
  auto& mp = r2tabmap<T>(ctx);
  int radix[4];
  constexpr bool F32 = sizeof(T) == 4;
  int ns = ymode ? r2r2_query<true, F32>(n, var, radix) : r2r2_query<false, F32>(n, var, radix);
  if (ns < 1 && var != 0) {
    var = 0;
    ns = ymode ? r2r2_query<true, F32>(n, var, radix) : r2r2_query<false, F32>(n, var, radix);
  }
  const int key = (n * 4 + var) * 2 + (ymode ? 1 : 0);
  auto it = mp.find(key);
  if (it != mp.end()) { *out = &it->second; return 0; }
  if (ns < 1) { *out = nullptr; return 0; }
  R2Tables<T> t;
  const long double pi = 3.14159265358979323846264338327950288L;
  int Ns = n;
  for (int s = 0; s < ns; ++s) {
    const int Rs = radix[s], L = Ns / Rs;
    if (L > 1) {
      std::vector<Cx<T>> h((size_t)(Rs - 1) * L);
      for (int r = 1; r < Rs; ++r)
        for (int o = 0; o < L; ++o) {
          const long double a = -2.0L * pi * (long double)((long long)o * r) / Ns;
          h[(size_t)(r - 1) * L + o] = Cx<T>{(T)cosl(a), (T)sinl(a)};
        }
      CK(cudaMalloc(&t.tw[s], sizeof(Cx<T>) * h.size()));
      CK(upload(t.tw[s], h.data(), sizeof(Cx<T>) * h.size()));
    }
    Ns = L;
  }
  std::vector<Cx<T>> mk((size_t)n / 2 + 1);
  for (int k = 0; k <= n / 2; ++k) {
    const long double a = pi * k / (2.0L * n);
    mk[k] = Cx<T>{(T)cosl(a), (T)sinl(a)};
  }
  CK(cudaMalloc(&t.mak, sizeof(Cx<T>) * mk.size()));
  CK(upload(t.mak, mk.data(), sizeof(Cx<T>) * mk.size()));
  auto res = mp.emplace(key, t);
  *out = &res.first->second;
  return 0;
}

template <class T> static int get_tables(cansb200_ctx* ctx, int n, FftTables<T>** out) {
  auto& mp = tabmap<T>(ctx);
  auto it = mp.find(n);
  if (it != mp.end()) { *out = &it->second; return 0; }
  FftTables<T> t;
  t.H = make_host_plan(n, K_R2HC);  // tables do not depend on the kind
  if (t.H.fast) {
    const int M = t.H.M;
    std::vector<C2<T>> tw(M), twp(M / 2 + 1), mak(M + 1);
    for (int i = 0; i < M; ++i) tw[i] = {(T)t.H.tw_re[i], (T)t.H.tw_im[i]};
    for (int i = 0; i <= M / 2; ++i) twp[i] = {(T)t.H.twp_re[i], (T)t.H.twp_im[i]};
    for (int i = 0; i <= M; ++i) mak[i] = {(T)t.H.mak_re[i], (T)t.H.mak_im[i]};
    CK(cudaMalloc(&t.tw, sizeof(C2<T>) * (M > 0 ? M : 1)));
    CK(cudaMalloc(&t.twp, sizeof(C2<T>) * (M / 2 + 1)));
    CK(cudaMalloc(&t.mak, sizeof(C2<T>) * (M + 1)));
    CK(cudaMalloc(&t.rev, sizeof(uint16_t) * (M > 0 ? M : 1)));
    CK(upload(t.tw, tw.data(), sizeof(C2<T>) * M));
    CK(upload(t.twp, twp.data(), sizeof(C2<T>) * (M / 2 + 1)));
    CK(upload(t.mak, mak.data(), sizeof(C2<T>) * (M + 1)));
    CK(upload(t.rev, t.H.rev.data(), sizeof(uint16_t) * M));
  }
  auto res = mp.emplace(n, t);
  *out = &res.first->second;
  return 0;
}

template <class T> static int get_direct_tables(cansb200_ctx* ctx, int n, int kind, DirectTables<T>** out) {
  auto& mp = dtabmap<T>(ctx);
  const int Q = slow_Q(n, kind);
  auto it = mp.find(Q);
  if (it != mp.end()) { *out = &it->second; return 0; }
  DirectTables<T> t;
  t.Q = Q;
  std::vector<C2<T>> cs(2 * (size_t)Q);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (long long m = 0; m < 2LL * Q; ++m) cs[m] = {(T)cosl(pi * m / Q), (T)sinl(pi * m / Q)};
  CK(cudaMalloc(&t.cs, sizeof(C2<T>) * cs.size()));
  CK(upload(t.cs, cs.data(), sizeof(C2<T>) * cs.size()));
  auto res = mp.emplace(Q, t);
  *out = &res.first->second;
  return 0;
}

// ---------------------------------------------------------------------------
// one batched r2r launch.  ymode: lines are `lines_per_group` consecutive
// elements (stride ls) and the transform runs with element stride es.
struct R2RGeom {
  long long in_es, out_es, in_ls, out_ls, in_gs, out_gs;
  int lines_per_group, ngroups, line_len, ymode;
  const void* row_tab = nullptr;   // distributed y transforms: peer-mapped output (forward) / input (backward) rows
  int smem_pad = 0;                // ... extra shared memory per CTA (caps the CTAs per SM while the kernel shares the GPU)
  int x0 = 0, g0 = 0;              // ... launched on a window of the slab: its first column / plane (the row table addresses whole rows)
};

// default plan per (length, mode, kind): variant 0 of r2r2_inst.cuh except where a sweep on B200 found better
// (scripts/bench_stages.py; FP64 y-512: the 16-samples-per-thread plan wins for the forward kinds and the
// cosine / sine inverses, the 8-samples-per-thread plan for HC2R)
static int r2_auto_variant(int n, int ymode, int kind, bool fp32) {
  if (!fp32 && ymode && n == 512 && kind != K_HC2R) return 3;
  return 0;
}

template <class T>
static int run_r2r(cansb200_ctx* ctx, int kind, int nt, const T* in, T* out, const R2RGeom& g, int tile_hint,
                   cudaStream_t st) {
  if (nt < 1 || g.lines_per_group < 1 || g.ngroups < 1) return 0;
  NvtxRange nvtx_r2r(g.ymode ? (kind_is_forward(kind) ? "fft_y_fwd" : "fft_y_bwd") : (kind_is_forward(kind) ? "fft_x_fwd" : "fft_x_bwd"));
  // fast path: two-for-one register transforms (r2r2.cuh) for the instantiated lengths
  if (!ctx->force_generic && kind_is_fast(kind)) {
    bool ok = true;
    if (g.ymode) {
      // column pairs are moved as 16-byte (2 x T) vectors
      const size_t al = 2 * sizeof(T);
      ok = g.in_ls == 1 && g.out_ls == 1 && (g.in_es % 2) == 0 && (g.out_es % 2) == 0 && (g.in_gs % 2) == 0 &&
           (g.out_gs % 2) == 0 && ((uintptr_t)in % al) == 0 && ((uintptr_t)out % al) == 0;
    } else {
      ok = g.in_es == 1 && g.out_es == 1;
    }
    R2Tables<T>* rt = nullptr;
    int var = ctx->r2_variant[g.ymode ? 1 : 0];
    if (var < 0) var = r2_auto_variant(nt, g.ymode, kind, sizeof(T) == 4);
    // cansb200_solve_fillps: the forward x transform of that solve reads u, v, w instead of p (variant 0 kernels)
    const bool fused = !g.ymode && ctx->fuse && kind_is_forward(kind);
    if (fused) var = 0;
    if (ok) {
      int rc = get_r2_tables<T>(ctx, nt, g.ymode, var, &rt);
      if (rc) return rc;
    }
    if (ok && rt) {
      R2Args<T> A;
      A.in = in; A.out = out;
      A.in_es = g.in_es; A.out_es = g.out_es; A.in_ls = g.in_ls; A.out_ls = g.out_ls; A.in_gs = g.in_gs; A.out_gs = g.out_gs;
      A.lines_per_group = g.lines_per_group; A.ngroups = g.ngroups; A.line_len = g.line_len; A.kind = kind;
      for (int s = 0; s < 4; ++s) A.tw[s] = rt->tw[s];
      A.mak = rt->mak;
      A.row_tab = (const R2Row<T>*)g.row_tab;
      A.x0 = g.x0; A.g0 = g.g0; A.smem_pad = g.smem_pad;
      A.flags = (ctx->r2_flags & 3) | ((!g.ymode && ctx->cur_xsplit) ? CB_R2_XSPLIT : 0);
      cb::g_r2_default_carveout = (ctx->r2_flags & 4) ? 0 : 1;
      if (fused) {
        const cansb200_ctx::FuseSrc& fs = *ctx->fuse;
        const long long delta = in - (const T*)fs.pin;   // this launch starts `delta / plane` planes into the slab
        if (delta < 0 || g.in_gs <= 0 || delta % g.in_gs != 0 || g.line_len != nt)
          return fail(CANSB200_EINVAL, "r2r: fused fillps source on a launch that is not a run of whole planes");
        R2Fill<T> F;
        F.u = (const T*)fs.u + delta; F.v = (const T*)fs.v + delta; F.w = (const T*)fs.w + delta;
        F.dzfi = (const T*)fs.dzfi;
        // fillps.f90:35-36: dtidxi = dti*dli(1), dtidyi = dti*dli(2), in the working precision
        F.dti = (T)fs.dti; F.dtidxi = (T)fs.dti * (T)fs.dxi; F.dtidyi = (T)fs.dti * (T)fs.dyi;
        F.sj = g.in_ls; F.sk = g.in_gs;
        F.k0 = (int)(delta / g.in_gs);
        F.any_rhsb = fs.any_rhsb;
        for (int d = 0; d < 3; ++d)
          for (int sd = 0; sd < 2; ++sd) { F.idx[d][sd] = fs.B.idx[d][sd]; F.val[d][sd] = (T)fs.B.val[d][sd]; }
        const int rcf = r2r2_run_fillps<T>(A, F, nt, st);
        if (rcf != 0) return fail(rcf < 0 ? CANSB200_ECUDA : CANSB200_EUNSUPPORTED, "r2r: fused fillps launch failed");
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
      }
      const int rc = !g.ymode ? r2r2_run<T, false, false>(A, nt, var, kind_is_forward(kind), st)
                     : g.row_tab ? r2r2_run<T, true, true>(A, nt, var, kind_is_forward(kind), st)
                                 : r2r2_run<T, true, false>(A, nt, var, kind_is_forward(kind), st);
      if (rc < 0) return fail(CANSB200_ECUDA, "r2r: fast-path launch failed");
      if (rc == 0) {
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
      }
    }
  }
  if (!g.ymode && ctx->fuse && kind_is_forward(kind))
    return fail(CANSB200_EUNSUPPORTED, "r2r: the fused fillps source needs a fast-path x length");
  if (g.row_tab) return fail(CANSB200_EUNSUPPORTED, "r2r: the distributed solve needs a fast-path y length (64..2048, 2^k or 3*2^k)");
  if (!g.ymode && ctx->cur_xsplit)
    return fail(CANSB200_EINVAL, "r2r: the plan keeps x in split order (pivot_dedup), which only the fast x transforms write; "
                                 "set CANSB200_CTX_FORCE_GENERIC before creating plans");
  FftTables<T>* tb;
  int rc = get_tables<T>(ctx, nt, &tb);
  if (rc) return rc;
  if (tb->H.fast && kind_is_fast(kind)) {
    FftArgs<T> A;
    A.P.n = nt; A.P.M = tb->H.M; A.P.kind = kind; A.P.nstages = (int)tb->H.radix.size();
    int rmax = 1;
    for (int s = 0; s < A.P.nstages; ++s) { A.P.radix[s] = tb->H.radix[s]; if (tb->H.radix[s] > rmax) rmax = tb->H.radix[s]; }
    A.P.tw = tb->tw; A.P.twp = tb->twp; A.P.mak = tb->mak; A.P.rev = tb->rev;
    A.in = in; A.out = out;
    A.in_es = g.in_es; A.out_es = g.out_es; A.in_ls = g.in_ls; A.out_ls = g.out_ls; A.in_gs = g.in_gs; A.out_gs = g.out_gs;
    A.lines_per_group = g.lines_per_group; A.ngroups = g.ngroups; A.line_len = g.line_len; A.ymode = g.ymode;
    const int nthr = 256;
    const size_t cap = 64 * 1024;
    if (g.ymode) {
      int cx = tile_hint > 0 ? tile_hint : 16;
      while (cx > 4 && (size_t)nt * cx * sizeof(T) > cap) cx /= 2;
      A.tile_lines = cx;
    } else {
      const size_t line_b = (size_t)LayX::line_len(A.P.M) * sizeof(T);
      long long nl = tile_hint > 0 ? tile_hint : ((long long)nthr * rmax + A.P.M - 1) / (A.P.M > 0 ? A.P.M : 1);
      if (nl * (long long)line_b > (long long)cap) nl = cap / line_b;
      if (nl > 1024) nl = 1024;
      if (nl < 1) nl = 1;
      A.tile_lines = (int)nl;
    }
    const size_t smem = tile_smem_elems(A) * sizeof(T);
    if (smem > 200 * 1024) return fail(CANSB200_EUNSUPPORTED, "r2r: line too long for the shared-memory tile");
    const long long ntiles = num_tiles(A);
    const unsigned grid = (unsigned)(ntiles < 0x7fffffffLL ? ntiles : 0x7fffffffLL);
    if (g.ymode) {
      auto kfn = fft_tile_kernel<T, LayY>;
      CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      kfn<<<grid, nthr, smem, st>>>(A, ntiles);
    } else {
      auto kfn = fft_tile_kernel<T, LayX>;
      CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      kfn<<<grid, nthr, smem, st>>>(A, ntiles);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  // direct O(n^2) evaluation
  DirectTables<T>* dt;
  rc = get_direct_tables<T>(ctx, nt, kind, &dt);
  if (rc) return rc;
  DirectArgs<T> D;
  D.n = nt; D.kind = kind; D.Q = dt->Q; D.cs = dt->cs; D.in = in; D.out = out;
  D.in_es = g.in_es; D.out_es = g.out_es; D.in_ls = g.in_ls; D.out_ls = g.out_ls; D.in_gs = g.in_gs; D.out_gs = g.out_gs;
  D.lines_per_group = g.lines_per_group; D.ngroups = g.ngroups; D.line_len = g.line_len;
  const size_t smem = (size_t)g.line_len * sizeof(T);
  if (smem > 200 * 1024) return fail(CANSB200_EUNSUPPORTED, "r2r(direct): line too long");
  auto kfn = r2r_direct_kernel<T>;
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const long long nlines = (long long)g.lines_per_group * g.ngroups;
  kfn<<<(unsigned)nlines, 128, smem, st>>>(D);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// thomas
template <class T> static ThomasDev<T> make_thomas(const cansb200_plan* pl, int nx, int ny, long long sj, long long sk,
                                                   int n_rows, int periodic, const T* lam, const T* a, const T* b, const T* c) {
  ThomasDev<T> D;
  D.nx = nx; D.ny = ny; D.n = n_rows; D.periodic = periodic; D.nn = periodic ? n_rows - 1 : n_rows;
  D.sj = sj; D.sk = sk; D.a = a; D.b = b; D.c = c; D.lam = lam; D.lam_sj = nx;
  D.m = pl->th_m; D.chunk_layout = 2; D.xb = 0; D.xn = nx; D.out_rows = nullptr; D.nopin = 0;
  D.dx = pl->dx; D.dy = pl->dy; D.nxu = pl->nxu; D.nyu = pl->nyu;
  D.zsj = (long long)D.nn * D.nxu; D.zsk = D.nxu;
  D.dt_mode = 0; D.dt_z1 = nullptr; D.dt_rp = nullptr; D.dt_slot_small = 0;
  D.jb = 1;
  // tall tiles: only where (j, k) rows form one uniform run -- the z-major field of the one-GPU solve
  if (pl->jb > 1 && sj == (long long)D.nn * sk && !periodic && n_rows == pl->th_n && (ny % pl->jb) == 0) {
    D.jb = pl->jb;
    D.m = pl->th_m_b;
  }
  return D;
}

static bool thomas_is_pipelined(const cansb200_plan* pl) { return pl->th_variant == 1 || pl->th_variant == 3; }

// ---- TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point; no libcuda link) ----
typedef CUresult (*cb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cb_encode_tiled_fn get_encode_tiled() {
  static cb_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (cb_encode_tiled_fn)q;
  }
  return fn;
}
// tiles of {box_cols columns (one 128-byte row segment) x box_rows rows}: rank 3 = (x, y, row), rank 4 adds the cache slot
static bool encode_tile_map(CUtensorMap* m, const void* base, size_t esz, int rank, const cuuint64_t dims[4],
                            const cuuint64_t strides_bytes[3], int box_rows, int box_cols, bool rows_second) {
  cb_encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint32_t box[4] = {(cuuint32_t)box_cols, 1u, (cuuint32_t)box_rows, 1u};   // field: (x, y, row)
  if (rows_second) { box[1] = (cuuint32_t)box_rows; box[2] = 1u; }             // pivot cache: (x, row, y, slot)
  cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  for (int d = 0; d + 1 < rank; ++d)
    if (strides_bytes[d] % 16 != 0 || strides_bytes[d] >= (1ULL << 40)) return false;
  if (((uintptr_t)base % 16) != 0) return false;
  const CUtensorMapDataType dt = esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  return enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class T>
static bool thomas_tma_maps(cansb200_plan* pl, const ThomasDev<T>& D, const T* p, int box_rows, int box_cols,
                            const CUtensorMap** mp, const CUtensorMap** mz) {
  if (!pl->use_tma || D.nx < box_cols) return false;
  const size_t esz = sizeof(T);
  const void* zb = pl->zsrc_base ? pl->zsrc_base : pl->zcache.p;
  const long long zslot = pl->zsrc_base ? pl->zsrc_slot : pl->slot_z;
  const bool tall = D.jb > 1;   // (j, k) rows as one run: (x, j nn + k, 0, slot) and (x, 0, j nn + k)
  const unsigned long long zkey[4] = {(unsigned long long)(uintptr_t)zb, ((unsigned long long)D.nxu << 32) | (unsigned long long)D.nyu,
                                      ((unsigned long long)D.nn << 20) | (unsigned long long)box_rows | (tall ? (1ULL << 60) : 0ULL),
                                      (unsigned long long)D.zsk * 1000003ULL + (unsigned long long)D.zsj};
  if (memcmp(zkey, pl->map_z_key, sizeof(zkey)) != 0) {
    // pivots: tensor (x, row, y, slot) -- the cache z[slot][j][k][i], or the Z[slot][k][j][i] of the distributed TDMA
    const cuuint64_t dims[4] = {(cuuint64_t)D.nxu, (cuuint64_t)D.nn * (tall ? D.nyu : 1), (cuuint64_t)(tall ? 1 : D.nyu), (cuuint64_t)pl->nslots};
    const cuuint64_t st[3] = {(cuuint64_t)D.zsk * esz, (cuuint64_t)D.zsj * esz, (cuuint64_t)zslot * esz};
    if (!encode_tile_map(&pl->map_z, zb, esz, 4, dims, st, box_rows, box_cols, true)) { pl->use_tma = false; return false; }
    memcpy(pl->map_z_key, zkey, sizeof(zkey));
  }
  *mz = &pl->map_z;
  const int ny_key = tall ? -D.ny : D.ny;   // tall and per-row maps of the same field are different descriptors
  for (auto& e : pl->map_p)
    if (e.p == p && e.sj == D.sj && e.sk == D.sk && e.nx == D.nx && e.ny == ny_key && e.nn == D.nn && e.box_rows == box_rows) {
      *mp = &e.m;
      return true;
    }
  cansb200_plan::PMap e;
  e.p = p; e.sj = D.sj; e.sk = D.sk; e.nx = D.nx; e.ny = ny_key; e.nn = D.nn; e.box_rows = box_rows;
  const cuuint64_t dims[4] = {(cuuint64_t)D.nx, (cuuint64_t)(tall ? 1 : D.ny), (cuuint64_t)D.nn * (tall ? D.ny : 1), 1};
  const cuuint64_t st[3] = {(cuuint64_t)D.sj * esz, (cuuint64_t)D.sk * esz, 0};
  if (!encode_tile_map(&e.m, p, esz, 3, dims, st, box_rows, box_cols, false)) return false;
  if (pl->map_p.size() >= 16) pl->map_p.erase(pl->map_p.begin());
  pl->map_p.push_back(e);
  *mp = &pl->map_p.back().m;
  return true;
}

template <class T, int MMAX, bool EXACT, int LDM, int COLS, int CL, bool TALL = false>
static int launch_pipe(cansb200_ctx* ctx, const ThomasDev<T>& D, const cansb200_plan* pl, T* p, T norm, const CUtensorMap* mp,
                       const CUtensorMap* mz, int box_rows, cudaStream_t st) {
  auto kfn = thomas_pipe_kernel<T, MMAX, EXACT, LDM, COLS, CL, TALL>;
  const size_t smem = thomas_pipe_smem<T, MMAX, COLS>();
  // per device: the shared-memory opt-in and the cluster occupancy belong to the device the context lives on
  static std::map<int, int> per_dev;   // device -> max resident clusters (0 for CL = 1)
  int dev = 0;
  CK(cudaGetDevice(&dev));
  auto itd = per_dev.find(dev);
  int max_clusters = itd != per_dev.end() ? itd->second : 0;
  if (itd == per_dev.end()) {
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 1) {
      // how many CTA pairs can be resident at once (GPCs with an odd SM count leave one SM out)
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(CL * ctx->num_sms); q.blockDim = dim3(CB_TH_THREADS); q.dynamicSmemBytes = smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      CK(cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &q));
      if (max_clusters < 1) return fail(CANSB200_ECUDA, "gaussel: no room for a CTA cluster");
    }
    per_dev[dev] = max_clusters;
  }
  const long long tiles = (long long)((D.xn + COLS - 1) / COLS) * (D.ny / (D.jb > 1 ? D.jb : 1));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute la[1];
  if (CL > 1) {
    long long nc = tiles < max_clusters ? tiles : max_clusters;   // persistent: one cluster per SM pair
    if (ctx->cta_cap > 0 && nc > ctx->cta_cap / CL) nc = ctx->cta_cap / CL > 0 ? ctx->cta_cap / CL : 1;
    cfg.gridDim = dim3((unsigned)(nc * CL));
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = CL; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    cfg.attrs = la; cfg.numAttrs = 1;
  } else {
    long long nb = tiles < ctx->num_sms ? tiles : ctx->num_sms;   // persistent: one CTA per SM
    if (ctx->cta_cap > 0 && nb > ctx->cta_cap) nb = ctx->cta_cap;
    cfg.gridDim = dim3((unsigned)nb);
  }
  cfg.blockDim = dim3(CB_TH_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  static const CUtensorMap dummy = {};
  const T* zb = pl->zsrc_base ? (const T*)pl->zsrc_base : (const T*)pl->zcache.p;
  const long long zslot = pl->zsrc_base ? pl->zsrc_slot : pl->slot_z;
  CK(cudaLaunchKernelEx(&cfg, kfn, D, (const CacheState*)pl->state.p, zb, (const T*)pl->p2cache.p,
                        (const T*)pl->dencache.p, zslot, pl->slot_den, p, norm, mp ? *mp : dummy, mz ? *mz : dummy, box_rows));
  ctx->launches++;
  return 0;
}

template <class T, int MMAX, int COLS, int CL, bool TALL = false>
static int launch_pipe_sel(cansb200_ctx* ctx, const ThomasDev<T>& D, cansb200_plan* pl, T* p, T norm, bool exact, bool vec,
                           cudaStream_t st) {
  if (exact && vec && COLS * sizeof(T) == 128 && pl->th_variant != 3) {
    // TMA boxes of 256 / 128 / 64 rows, whichever divides the rows of one CTA
    const int rows = (CB_TH_THREADS / COLS) * MMAX;
    const int box_rows = rows % 256 == 0 ? 256 : (rows % 128 == 0 ? 128 : 64);
    const CUtensorMap *mp = nullptr, *mz = nullptr;
    if (thomas_tma_maps<T>(pl, D, p, box_rows, COLS, &mp, &mz))
      return launch_pipe<T, MMAX, true, CB_TH_LD_TMA, COLS, CL, TALL>(ctx, D, pl, p, norm, mp, mz, box_rows, st);
  }
  if (exact && vec) return launch_pipe<T, MMAX, true, CB_TH_LD_VEC, COLS, CL, TALL>(ctx, D, pl, p, norm, nullptr, nullptr, 0, st);
  if (vec) return launch_pipe<T, MMAX, false, CB_TH_LD_VEC, COLS, CL, TALL>(ctx, D, pl, p, norm, nullptr, nullptr, 0, st);
  return launch_pipe<T, MMAX, false, CB_TH_LD_ELEM, COLS, CL, TALL>(ctx, D, pl, p, norm, nullptr, nullptr, 0, st);
}

// sizes of the pivot cache for the plan's deduplication flags
static void plan_cache_extents(cansb200_plan* pl) {
  cansb200_ctx* ctx = pl->ctx;
  const int nx = ctx->n_z[0], ny = ctx->n_z[1];
  pl->nxu = pl->dx ? nx / 2 + pl->th_cols : nx;
  pl->nyu = pl->dy ? ny / 2 + 1 : ny;
  pl->slot_z = (long long)pl->nxu * pl->nyu * pl->th_nn;
  pl->slot_den = (long long)pl->nxu * pl->nyu;
}
static int plan_cache_alloc(cansb200_plan* pl) {
  cansb200_ctx* ctx = pl->ctx;
  plan_cache_extents(pl);
  if (!ctx->dtdma && pl->zcache.ensure((size_t)pl->slot_z * pl->nslots * ctx->esz)) return fail(CANSB200_ENOMEM, "plan: pivot cache");
  if (pl->periodic_z && !ctx->dtdma) {
    if (pl->p2cache.ensure((size_t)pl->slot_z * pl->nslots * ctx->esz)) return fail(CANSB200_ENOMEM, "plan: p2 cache");
    if (pl->dencache.ensure((size_t)pl->slot_den * pl->nslots * ctx->esz)) return fail(CANSB200_ENOMEM, "plan: den cache");
  }
  if (pl->state.ensure(sizeof(CacheState))) return fail(CANSB200_ENOMEM, "plan: cache state");
  CacheState cs;
  memset(&cs, 0, sizeof(cs));
  cs.nslots = pl->nslots;
  CK(upload(pl->state.p, &cs, sizeof(cs)));
  return 0;
}

// pivot cache: content hash of (a, b, c, lambda) -> slot select -> factorisation on a miss
template <class T> static int gaussel_prepare(cansb200_plan* pl, ThomasDev<T>& D, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  for (int attempt = 0; attempt < 2; ++attempt) {
    CacheState* cs = (CacheState*)pl->state.p;
    const long long nlam = (long long)D.nx * D.ny;
    const long long total = 3LL * D.n + nlam;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 592) blocks = 592;
    thomas_hash_kernel<T><<<blocks, 256, 0, st>>>(D, cs);
    ctx->launches++;
    if ((D.dx | D.dy) && !pl->sym_checked) {
      // The deduplicated cache assumes lambda(i) = lambda(nx - i) (it is, for the eigenvalues initsolver builds).  The hash
      // kernel checks every element on every solve; the FIRST solve of a plan reads the verdict back (one host sync, like
      // the first-use allocations) and, if the caller's lambdaxy is not symmetric, falls back to the full-size cache.
      // (page-locked destination: a pageable one may synchronise the whole device, and other streams may hold device-side waits)
      CK(cudaMemcpyAsync(ctx->pinned_word, &cs->sym_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const int bad = *(volatile int*)ctx->pinned_word;
      pl->sym_checked = true;
      if (bad) {
        pl->dx = pl->dy = 0;
        pl->zcache.release(); pl->p2cache.release(); pl->dencache.release();
        const int rc = plan_cache_alloc(pl);
        if (rc) return rc;
        memset(pl->map_z_key, 0, sizeof(pl->map_z_key));
        D.dx = D.dy = 0; D.nxu = pl->nxu; D.nyu = pl->nyu;
        D.zsj = (long long)D.nn * D.nxu; D.zsk = D.nxu;
        continue;   // hash again without the flags
      }
    }
    thomas_select_kernel<<<1, 32, 0, st>>>(cs);
    const long long ncol_s = (long long)(D.dx ? D.nxu : D.nx) * (D.dy ? D.ny / 2 + 1 : D.ny);   // stored columns
    thomas_factor_kernel<T><<<(unsigned)((ncol_s + 127) / 128), 128, 0, st>>>(D, cs, (T*)pl->zcache.p, (T*)pl->p2cache.p,
                                                                              (T*)pl->dencache.p, pl->slot_z, pl->slot_den);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return 0;
  }
  return fail(CANSB200_ECUDA, "gaussel: pivot cache set-up failed");
}

// substitution on the column window D.xb .. D.xb + D.xn - 1
template <class T> static int gaussel_apply(cansb200_plan* pl, const ThomasDev<T>& D, T* p, T norm, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  NvtxRange nvtx_th("gaussel");
  CacheState* cs = (CacheState*)pl->state.p;
  if (thomas_is_pipelined(pl)) {
    // 16-byte tile copies need every row segment aligned: even (FP64) / multiple-of-4 (FP32) column counts and offsets
    const long long vw = 16 / (long long)sizeof(T);
    const bool vec = (D.nx % vw) == 0 && (D.xb % vw) == 0 && (D.sk % vw) == 0 && (D.sj % vw) == 0 && ((uintptr_t)p % 16) == 0 &&
                     ((uintptr_t)(pl->zsrc_base ? pl->zsrc_base : pl->zcache.p) % 16) == 0 &&
                     ((pl->zsrc_base ? pl->zsrc_slot : pl->slot_z) % vw) == 0 && (D.xn % vw) == 0 && (D.zsk % vw) == 0 && (D.zsj % vw) == 0;
    const int mmax_eff = D.jb > 1 ? pl->th_mmax_b : pl->th_mmax;
    const bool exact = D.m == mmax_eff;
    if constexpr (sizeof(T) == 8) {
      if (pl->th_cols == 16 && pl->th_cl == 1 && D.jb > 1) {
        if (mmax_eff == 4) return launch_pipe_sel<T, 4, 16, 1, true>(ctx, D, pl, p, norm, exact, vec, st);
        if (mmax_eff == 8) return launch_pipe_sel<T, 8, 16, 1, true>(ctx, D, pl, p, norm, exact, vec, st);
      } else if (pl->th_cols == 16 && pl->th_cl == 1) {
        if (mmax_eff == 4) return launch_pipe_sel<T, 4, 16, 1>(ctx, D, pl, p, norm, exact, vec, st);
        if (mmax_eff == 8) return launch_pipe_sel<T, 8, 16, 1>(ctx, D, pl, p, norm, exact, vec, st);
      } else if (pl->th_cols == 16 && pl->th_cl == 2) {
        if (pl->th_mmax == 6) return launch_pipe_sel<T, 6, 16, 2>(ctx, D, pl, p, norm, exact, vec, st);
        if (pl->th_mmax == 8) return launch_pipe_sel<T, 8, 16, 2>(ctx, D, pl, p, norm, exact, vec, st);
      } else if (pl->th_cols == 8) {
        if (pl->th_mmax == 6) return launch_pipe_sel<T, 6, 8, 1>(ctx, D, pl, p, norm, exact, vec, st);
        if (pl->th_mmax == 8) return launch_pipe_sel<T, 8, 8, 1>(ctx, D, pl, p, norm, exact, vec, st);
      }
    } else {
      // FP32: 32 columns make the 128-byte row segment (32 chunks of up to 16 rows per column)
      if (pl->th_cols == 32 && pl->th_cl == 1 && D.jb > 1) {
        if (mmax_eff == 8) return launch_pipe_sel<T, 8, 32, 1, true>(ctx, D, pl, p, norm, exact, vec, st);
        if (mmax_eff == 16) return launch_pipe_sel<T, 16, 32, 1, true>(ctx, D, pl, p, norm, exact, vec, st);
      } else if (pl->th_cols == 32 && pl->th_cl == 1) {
        if (mmax_eff == 8) return launch_pipe_sel<T, 8, 32, 1>(ctx, D, pl, p, norm, exact, vec, st);
        if (mmax_eff == 16) return launch_pipe_sel<T, 16, 32, 1>(ctx, D, pl, p, norm, exact, vec, st);
      } else if (pl->th_cols == 32 && pl->th_cl == 2) {
        if (pl->th_mmax == 12) return launch_pipe_sel<T, 12, 32, 2>(ctx, D, pl, p, norm, exact, vec, st);
        if (pl->th_mmax == 16) return launch_pipe_sel<T, 16, 32, 2>(ctx, D, pl, p, norm, exact, vec, st);
      } else if (pl->th_cols == 16) {
        if (pl->th_mmax == 16) return launch_pipe_sel<T, 16, 16, 1>(ctx, D, pl, p, norm, exact, vec, st);
      }
    }
    return fail(CANSB200_EUNSUPPORTED, "gaussel: no pipelined kernel for this chunk length");
  }
  if (D.xb != 0 || D.xn != D.nx) return fail(CANSB200_EUNSUPPORTED, "gaussel: the sequential variant has no column windows");
  const long long ncol = (long long)D.nx * D.ny;
  thomas_seq_kernel<T><<<(unsigned)((ncol + 127) / 128), 128, 0, st>>>(D, cs, (const T*)pl->zcache.p, (const T*)pl->p2cache.p,
                                                                     (const T*)pl->dencache.p, pl->slot_z, pl->slot_den, p, norm);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <class T>
static int run_gaussel(cansb200_plan* pl, T* p, int nx, int ny, long long sj, long long sk, int n_rows, int periodic, T norm,
                       const T* lam, const T* a, const T* b, const T* c, cudaStream_t st) {
  if (n_rows != pl->th_n || periodic != pl->periodic_z)
    return fail(CANSB200_EINVAL, "gaussel: n_rows / periodicity differ from the plan's");
  if (pl->th_nn < 1) return fail(CANSB200_EINVAL, "gaussel: empty system");
  ThomasDev<T> D = make_thomas<T>(pl, nx, ny, sj, sk, n_rows, periodic, lam, a, b, c);
  int rc = gaussel_prepare<T>(pl, D, st);
  if (rc) return rc;
  prof_mark(pl->ctx, st);
  return gaussel_apply<T>(pl, D, p, norm, st);
}

// ---------------------------------------------------------------------------
static void find_fft(char b0, char b1, char cf, int& kf, int& kb, double& n1, double& n2) {
  // src/fft.f90:260-313
  const std::string key{b0, b1};
  n1 = 2.0; n2 = 0.0;
  if (key == "PP") { kf = K_R2HC; kb = K_HC2R; n1 = 1.0; return; }
  if (cf == 'c') {
    if (key == "NN") { kf = K_REDFT10; kb = K_REDFT01; }
    else if (key == "DD") { kf = K_RODFT10; kb = K_RODFT01; }
    else if (key == "ND") { kf = K_REDFT11; kb = K_REDFT11; }
    else { kf = K_RODFT11; kb = K_RODFT11; }
  } else {
    if (key == "NN") { kf = K_REDFT00; kb = K_REDFT00; n2 = -1.0; }
    else if (key == "DD") { kf = K_RODFT00; kb = K_RODFT00; n2 = 1.0; }
    else if (key == "ND") { kf = K_REDFT10; kb = K_REDFT01; }
    else { kf = K_RODFT01; kb = K_RODFT10; }
  }
}

extern "C" {

const char* cansb200_last_error(void) { return g_err.c_str(); }
int cansb200_version(void) { return 100; }

static void split_starts(int n, int P, std::vector<int>& st) {
  // first n mod P ranks get one extra point (dependencies/2decomp-fft/src/decomp_2d.f90:1018-1029,
  // dependencies/cuDecomp/src/cudecomp.cc:1348-1357)
  st.assign(P + 1, 0);
  const int base = n / P, rem = n % P;
  for (int r = 0; r < P; ++r) st[r + 1] = st[r] + base + (r < rem ? 1 : 0);
}

struct DistBlob {
  cudaIpcMemHandle_t handle;
  int rank, device;
  unsigned long long bytes;
};

int cansb200_init(cansb200_ctx** out, const int ng[3], const int dims[2], int ipencil_axis, int rank, int nranks,
                  const void* nccl_id, int is_fp32) {
  (void)nccl_id;   // the exchange runs over CUDA IPC peer mappings (cansb200_dist_export / _connect), not NCCL
  if (!out || !ng || !dims) return fail(CANSB200_EINVAL, "init: null argument");
  if (ng[0] < 1 || ng[1] < 1 || ng[2] < 1) return fail(CANSB200_EINVAL, "init: ng must be positive");
  if (ipencil_axis < 1 || ipencil_axis > 3) return fail(CANSB200_EINVAL, "init: ipencil_axis must be 1, 2 or 3");
  // The local array of `solver` is the pencil along ipencil_axis (src/initmpi.f90:198-260).  With dims = [1, P] the x and
  // the y pencil are the same z slab (nx, ny, nz/P), and on one rank every pencil is the whole grid; z pencils of a
  // decomposed grid (y split) would enter the solve at the tridiagonal stage's layout and are not implemented.
  if (ipencil_axis == 3 && nranks > 1)
    return fail(CANSB200_EUNSUPPORTED, "init: ipencil_axis = 3 on a decomposed grid is not implemented (use 1 or 2 with dims = [1, P])");
  if (nranks < 1 || nranks > CB_MAX_RANKS || rank < 0 || rank >= nranks) return fail(CANSB200_EINVAL, "init: bad rank / nranks");
  if (dims[0] != 1 || dims[1] != nranks)
    return fail(CANSB200_EUNSUPPORTED, "init: only dims = [1, nranks] (z slabs) is implemented");
  if (nranks > ng[1] || nranks > ng[2]) return fail(CANSB200_EINVAL, "init: more ranks than y or z planes");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail(CANSB200_ECUDA, "init: no CUDA device");
  auto* c = new cansb200_ctx();
  split_starts(ng[1], nranks, c->ys);
  split_starts(ng[2], nranks, c->zs);
  for (int d = 0; d < 3; ++d) { c->ng[d] = ng[d]; c->n[d] = ng[d]; c->lo[d] = 1; c->n_z[d] = ng[d]; c->lo_z[d] = 1; }
  c->n[2] = c->zs[rank + 1] - c->zs[rank];   c->lo[2] = c->zs[rank] + 1;      // x pencil: (nx, ny, nz/P)
  c->n_z[1] = c->ys[rank + 1] - c->ys[rank]; c->lo_z[1] = c->ys[rank] + 1;    // z pencil: (nx, ny/P, nz)
  c->dims[0] = dims[0]; c->dims[1] = dims[1]; c->ipencil_axis = ipencil_axis; c->rank = rank; c->nranks = nranks;
  c->is_fp32 = is_fp32 ? 1 : 0;
  c->esz = is_fp32 ? 4 : 8;
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      c->num_sms = sms;
  }
  const size_t nel = (size_t)c->n[0] * c->n[1] * c->n[2];
  if (c->scratch.ensure(nel * c->esz)) { delete c; return fail(CANSB200_ENOMEM, "init: scratch allocation failed"); }
  if (cudaHostAlloc((void**)&c->pinned_word, 64, cudaHostAllocDefault) != cudaSuccess) {
    c->scratch.release(); delete c;
    return fail(CANSB200_ENOMEM, "init: pinned word");
  }
  if (nranks > 1) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t nel_z = (size_t)c->n_z[0] * c->n_z[1] * c->n_z[2];
    c->off_C = 4096;
    c->off_XB = c->off_C + up(nel_z * c->esz);
    c->region_bytes = c->off_XB + up(nel * c->esz);
    if (cudaMalloc(&c->region, c->region_bytes) != cudaSuccess) { c->scratch.release(); delete c; return fail(CANSB200_ENOMEM, "init: exchange region"); }
    clear_dev(c->region, 4096);
    // host-mapped: the waiting kernels set it on a time-out, the host reads it without synchronising
    if (cudaHostAlloc((void**)&c->dist_status, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
      cudaFree(c->region); c->scratch.release(); delete c;
      return fail(CANSB200_ENOMEM, "init: status word");
    }
    *c->dist_status = 0;
    if (cudaMalloc((void**)&c->dist_status_dev, sizeof(int)) != cudaSuccess) {
      cudaFreeHost(c->dist_status); cudaFree(c->region); c->scratch.release(); delete c;
      return fail(CANSB200_ENOMEM, "init: status word");
    }
    clear_dev(c->dist_status_dev, sizeof(int));
    c->peer.assign(nranks, nullptr);
    c->peer[rank] = c->region;
  }
  *out = c;
  return 0;
}

int cansb200_dist_blob_size(void) { return (int)sizeof(DistBlob); }

int cansb200_dist_export(cansb200_ctx* c, void* blob) {
  if (!c || !blob) return fail(CANSB200_EINVAL, "dist_export: null argument");
  if (c->nranks < 2) return fail(CANSB200_EINVAL, "dist_export: single-rank context");
  DistBlob b;
  memset(&b, 0, sizeof(b));
  CK(cudaIpcGetMemHandle(&b.handle, c->region));
  CK(cudaGetDevice(&b.device));
  b.rank = c->rank;
  b.bytes = c->region_bytes;
  memcpy(blob, &b, sizeof(b));
  return 0;
}

}  // extern "C"

// device-side row tables of the peer-mapped stores / loads.  `packed`: the rows of a periodic y direction are dealt out
// to the z pencils by their position in the _OPENACC eigenvalue order (pack_index), so that the local slice of
// lambdaxy an unchanged OpenACC-build initsolver hands over matches the rows a rank owns.
template <class T> static int build_dist_tables(cansb200_ctx* c, bool packed) {
  const int P = c->nranks, r = c->rank;
  const long long nx = c->ng[0];
  const int ny = c->ng[1], nz = c->ng[2];
  const long long nzl_r = c->zs[r + 1] - c->zs[r];
  std::vector<R2Row<T>> yf(ny), yb(ny);
  T* XB_r = (T*)((char*)c->region + c->off_XB);
  for (int j = 0; j < ny; ++j) {
    const int pj = packed ? pack_index(j, ny) : j;   // position that decides the owner and the local row
    int s = 0;
    while (pj >= c->ys[s + 1]) ++s;
    const long long nyl_s = c->ys[s + 1] - c->ys[s], jl = pj - c->ys[s];
    T* C_s = (T*)((char*)c->peer[s] + c->off_C);        // C sits at the same offset in every rank's region
    // forward y output row j of my plane g: z pencil of rank s, plane zs[r] + g, row jl
    yf[j].ptr = C_s + ((long long)c->zs[r] * nyl_s + jl) * nx;
    yf[j].gs = nyl_s * nx;
    // backward y input row j of my plane g: block s of my way-back buffer [s][jl][g][i]
    yb[j].ptr = XB_r + (nzl_r * c->ys[s] + jl * nzl_r) * nx;
    yb[j].gs = nx;
  }
  DevBuf& tf = packed ? c->ytab_fwd_pk : c->ytab_fwd;
  DevBuf& tb = packed ? c->ytab_bwd_pk : c->ytab_bwd;
  if (tf.ensure(sizeof(R2Row<T>) * ny) || tb.ensure(sizeof(R2Row<T>) * ny)) return fail(CANSB200_ENOMEM, "dist_connect: tables");
  CK(upload(tf.p, yf.data(), sizeof(R2Row<T>) * ny));
  CK(upload(tb.p, yb.data(), sizeof(R2Row<T>) * ny));
  if (packed) { c->tabs_pk = true; return 0; }
  std::vector<OutRow<T>> zt(nz);
  for (int s = 0; s < P; ++s) {
    const long long nzl_s = c->zs[s + 1] - c->zs[s];
    T* XB_s = (T*)((char*)c->peer[s] + c->peer_off_XB[s]);
    for (int k = c->zs[s]; k < c->zs[s + 1]; ++k) {   // result row k of my columns: block r of rank s's way-back buffer, [j][k][i]
      zt[k].ptr = XB_s + nzl_s * c->ys[r] * nx + (long long)(k - c->zs[s]) * nx;
      zt[k].sj = nzl_s * nx;
    }
  }
  if (c->ztab.ensure(sizeof(OutRow<T>) * nz)) return fail(CANSB200_ENOMEM, "dist_connect: tables");
  CK(upload(c->ztab.p, zt.data(), sizeof(OutRow<T>) * nz));
  return 0;
}

extern "C" {

int cansb200_dist_connect(cansb200_ctx* c, const void* blobs) {
  if (!c || !blobs) return fail(CANSB200_EINVAL, "dist_connect: null argument");
  if (c->nranks < 2) return fail(CANSB200_EINVAL, "dist_connect: single-rank context");
  const DistBlob* b = (const DistBlob*)blobs;
  c->peer_off_XB.assign(c->nranks, 0);
  for (int s = 0; s < c->nranks; ++s) {
    if (b[s].rank != s) return fail(CANSB200_EINVAL, "dist_connect: blobs must be ordered by rank");
    // offsets inside rank s's region follow from the global splits
    const size_t nel_z = (size_t)c->ng[0] * (c->ys[s + 1] - c->ys[s]) * c->ng[2];
    c->peer_off_XB[s] = 4096 + ((nel_z * c->esz + 255) & ~(size_t)255);
    if (s == c->rank) continue;
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, b[s].handle, cudaIpcMemLazyEnablePeerAccess));
    c->peer[s] = q;
  }
  int rc = c->is_fp32 ? build_dist_tables<float>(c, false) : build_dist_tables<double>(c, false);
  if (rc) return rc;
  rc = c->is_fp32 ? build_dist_tables<float>(c, true) : build_dist_tables<double>(c, true);
  if (rc) return rc;
  c->connected = true;
  return 0;
}

int cansb200_dist_connect_local(cansb200_ctx* const* ctxs, int n) {
  // all ranks of the decomposition live in THIS process on THIS device (one context each, solves issued on one
  // stream per rank): the peers' regions are ordinary device pointers.  Same kernels, same row tables, same flags
  // as the CUDA-IPC rendezvous -- what a single-GPU box can test of the multi-GPU path.
  if (!ctxs || n < 2) return fail(CANSB200_EINVAL, "dist_connect_local: need at least two contexts");
  for (int r = 0; r < n; ++r) {
    cansb200_ctx* c = ctxs[r];
    if (!c || c->nranks != n || c->rank != r) return fail(CANSB200_EINVAL, "dist_connect_local: contexts must be ordered by rank, nranks = n");
    if (c->connected) return fail(CANSB200_EINVAL, "dist_connect_local: context already connected");
    for (int d = 0; d < 3; ++d)
      if (c->ng[d] != ctxs[0]->ng[d] || c->is_fp32 != ctxs[0]->is_fp32) return fail(CANSB200_EINVAL, "dist_connect_local: contexts differ");
  }
  for (int r = 0; r < n; ++r) {
    cansb200_ctx* c = ctxs[r];
    c->peer_off_XB.assign(n, 0);
    for (int s = 0; s < n; ++s) {
      c->peer_off_XB[s] = ctxs[s]->off_XB;
      c->peer[s] = ctxs[s]->region;
    }
    c->local_peers = true;
    int rc = c->is_fp32 ? build_dist_tables<float>(c, false) : build_dist_tables<double>(c, false);
    if (rc) return rc;
    rc = c->is_fp32 ? build_dist_tables<float>(c, true) : build_dist_tables<double>(c, true);
    if (rc) return rc;
    c->connected = true;
  }
  return 0;
}

int cansb200_dist_status(cansb200_ctx* c, int* status) {
  if (!c || !status) return fail(CANSB200_EINVAL, "dist_status: null argument");
  *status = c->dist_status ? *(volatile int*)c->dist_status : 0;
  return 0;
}

int cansb200_finalize(cansb200_ctx* c) {
  if (!c) return 0;
  c->scratch.release(); c->scratch2.release(); c->work2.release(); c->staging.release(); c->coef.release(); c->zero_lam.release();
  for (auto& kv : c->tabs64) { cudaFree(kv.second.tw); cudaFree(kv.second.twp); cudaFree(kv.second.mak); cudaFree(kv.second.rev); }
  for (auto& kv : c->tabs32) { cudaFree(kv.second.tw); cudaFree(kv.second.twp); cudaFree(kv.second.mak); cudaFree(kv.second.rev); }
  for (int q = 0; q < (int)c->peer.size(); ++q)
    if (q != c->rank && c->peer[q] && !c->local_peers) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->region) cudaFree(c->region);
  if (c->dist_status) cudaFreeHost(c->dist_status);
  if (c->pinned_word) cudaFreeHost(c->pinned_word);
  if (c->dist_status_dev) cudaFree(c->dist_status_dev);
  c->sendb.release(); c->ytab_fwd_loc.release(); c->ytab_fwd_loc_pk.release(); c->ztab_loc.release();
  for (cudaStream_t q : c->dist_cs) if (q) cudaStreamDestroy(q);
  if (c->last_done) cudaEventDestroy(c->last_done);
  if (c->dist_sF) cudaStreamDestroy(c->dist_sF);
  if (c->dist_sT) cudaStreamDestroy(c->dist_sT);
  if (c->dist_sB) cudaStreamDestroy(c->dist_sB);
  for (cudaEvent_t q : c->dist_ev) cudaEventDestroy(q);
  for (cudaEvent_t q : c->prof_events) cudaEventDestroy(q);
  c->ytab_fwd.release(); c->ytab_bwd.release(); c->ztab.release(); c->ytab_fwd_pk.release(); c->ytab_bwd_pk.release();
  c->lam_perm.release();
  { auto zp = c->zplans; c->zplans.clear(); for (auto& kv : zp) cansb200_plan_destroy(kv.second); }
  for (auto& rec : c->fftini_recs) { for (auto& kv : rec.byz) cansb200_plan_destroy(kv.second); rec.byz.clear(); }
  for (cudaStream_t q : c->aux) cudaStreamDestroy(q);
  for (cudaEvent_t q : c->aux_done) cudaEventDestroy(q);
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  for (auto& q : c->pinned) cudaHostUnregister(q.first);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (cudaEvent_t q : c->chunk_ev) cudaEventDestroy(q);
  for (auto& kv : c->r2tabs64) { for (int q = 0; q < 4; ++q) cudaFree(kv.second.tw[q]); cudaFree(kv.second.mak); }
  for (auto& kv : c->r2tabs32) { for (int q = 0; q < 4; ++q) cudaFree(kv.second.tw[q]); cudaFree(kv.second.mak); }
  for (auto& kv : c->dtabs64) cudaFree(kv.second.cs);
  for (auto& kv : c->dtabs32) cudaFree(kv.second.cs);
  delete c;
  return 0;
}

int cansb200_get_extents(const cansb200_ctx* c, int n[3], int lo[3], int n_z[3], int lo_z[3]) {
  if (!c) return fail(CANSB200_EINVAL, "null ctx");
  for (int d = 0; d < 3; ++d) {
    if (n) n[d] = c->n[d];
    if (lo) lo[d] = c->lo[d];
    if (n_z) n_z[d] = c->n_z[d];
    if (lo_z) lo_z[d] = c->lo_z[d];
  }
  return 0;
}

int cansb200_ctx_set(cansb200_ctx* c, int what, int value) {
  if (!c) return fail(CANSB200_EINVAL, "null ctx");
  if (what == CANSB200_CTX_FORCE_GENERIC) { c->force_generic = value ? 1 : 0; return 0; }
  if (what == CANSB200_CTX_CHAIN_COLS) {
    if (value < -1 || (value > 0 && (value % 16) != 0)) return fail(CANSB200_EINVAL, "ctx_set: chain_cols must be -1 (auto), 0 (off) or a multiple of 16");
    c->chain_cols = value;
    return 0;
  }
  if (what == CANSB200_CTX_CHAIN_STREAMS) {
    if (value < 1 || value > 8) return fail(CANSB200_EINVAL, "ctx_set: chain_streams must be 1..8");
    c->chain_nstreams = value;
    return 0;
  }
  if (what == CANSB200_CTX_X_VARIANT || what == CANSB200_CTX_Y_VARIANT) {
    if (value < -1 || value > 3) return fail(CANSB200_EINVAL, "ctx_set: variant must be -1 (per-kind default) or 0..3");
    c->r2_variant[what == CANSB200_CTX_Y_VARIANT ? 1 : 0] = value;
    return 0;
  }
  if (what == CANSB200_CTX_DIST_WINDOWS) {
    if (value < -1 || value == 0 || value > CB_MAX_WINDOWS) return fail(CANSB200_EINVAL, "ctx_set: dist_windows must be -1 (auto) or 1..8");
    c->dist_windows = value;
    return 0;
  }
  if (what == CANSB200_CTX_DIST_MODE) {
    if (value < -1 || value > 2) return fail(CANSB200_EINVAL, "ctx_set: dist_mode must be -1 (auto), 0, 1 or 2");
    c->dist_mode = value;
    return 0;
  }
  if (what == CANSB200_CTX_DIST_CHUNKS) {
    if (value < -1 || value == 0 || value > CB_MAX_WINDOWS) return fail(CANSB200_EINVAL, "ctx_set: dist_chunks must be -1 (auto) or 1..8");
    c->dist_chunks = value;
    return 0;
  }
  if (what == CANSB200_CTX_DIST_SPLIT_PAD) {
    if (value < -1 || value > 160) return fail(CANSB200_EINVAL, "ctx_set: dist_split_pad must be -1 (auto) or 0..160 KB");
    c->dist_split_pad = value;
    return 0;
  }
  if (what == CANSB200_CTX_DIST_THOMAS_CTAS) {
    if (value < -1 || value == 0) return fail(CANSB200_EINVAL, "ctx_set: dist_thomas_ctas must be -1 (auto) or positive");
    c->dist_thomas_ctas = value;
    return 0;
  }
  if (what == CANSB200_CTX_DTDMA) {
    if (c->nplans > 0) return fail(CANSB200_EINVAL, "ctx_set: CANSB200_CTX_DTDMA must be set before any plan is created on the context");
    // as in the reference (src/solver.f90:43-48, src/initsolver.f90 with is_poisson_dtdma): the "z pencil" extents that size
    // lambdaxy and a, b, c become those of the y pencil = my slab: lambdaxy(nx, ny), a/b/c(nz_local)
    c->dtdma = (value && c->nranks > 1) ? 1 : 0;
    for (int d = 0; d < 3; ++d) { c->n_z[d] = c->ng[d]; c->lo_z[d] = 1; }
    if (c->dtdma) {
      c->n_z[2] = c->zs[c->rank + 1] - c->zs[c->rank];
      c->lo_z[2] = c->zs[c->rank] + 1;
    } else {
      c->n_z[1] = c->ys[c->rank + 1] - c->ys[c->rank];
      c->lo_z[1] = c->ys[c->rank] + 1;
    }
    return 0;
  }
  if (what == CANSB200_CTX_DTDMA_TILED) {
    if (value < -1 || value > 2) return fail(CANSB200_EINVAL, "ctx_set: dtdma_tiled must be -1 / 1 (auto), 0 (never) or 2 (always)");
    c->dtdma_tiled = value < 0 ? 1 : value;
    return 0;
  }
  if (what == CANSB200_CTX_ZMAJOR) {
    c->zmajor = value ? 1 : 0;
    return 0;
  }
  if (what == CANSB200_CTX_PIN_HOST) {
    c->pin_host = value ? 1 : 0;
    return 0;
  }
  if (what == CANSB200_CTX_HOST_CHUNKS) {
    if (value < 1 || value > 64) return fail(CANSB200_EINVAL, "ctx_set: host_chunks must be 1..64");
    c->host_chunks = value;
    return 0;
  }
  if (what == CANSB200_CTX_AUX_3D) {
    c->aux_3d = value ? 1 : 0;
    return 0;
  }
  if (what == CANSB200_CTX_FUSE_FILLPS) {
    c->fuse_fillps = value ? 1 : 0;
    return 0;
  }
  if (what == CANSB200_CTX_R2_FLAGS) {
    if (value < 0 || value > 7) return fail(CANSB200_EINVAL, "ctx_set: r2 flags must be 0..7");
    c->r2_flags = value;
    return 0;
  }
  return fail(CANSB200_EINVAL, "ctx_set: unknown switch");
}

int cansb200_get_work(cansb200_ctx* c, int which, void** ptr, size_t* nelem) {
  // src/rk.f90:27-29 aliases three pencil-sized device buffers (`work`, `solver_buf_0`, `solver_buf_1`) as scratch of
  // the momentum step; they are only live between solver calls, so the library lends its own scratch:
  // 0 = A (x pencil), 1 = B (z-major intermediate), 2 = a third pencil.  Any solve on the context clobbers them.
  if (!c || !ptr) return fail(CANSB200_EINVAL, "null argument");
  if (which < 0 || which > 2) return fail(CANSB200_EINVAL, "get_work: which must be 0, 1 or 2");
  const size_t nel = (size_t)c->n[0] * c->n[1] * c->n[2];
  DevBuf* b = which == 0 ? &c->scratch : (which == 1 ? &c->scratch2 : &c->work2);
  if (b->ensure(nel * c->esz)) return fail(CANSB200_ENOMEM, "get_work: allocation failed");
  *ptr = b->p;
  if (nelem) *nelem = b->bytes / c->esz;
  return 0;
}

int cansb200_plan_create(cansb200_ctx* ctx, cansb200_plan** out, const char bc[6], const char cf[3],
                         const cansb200_options* opt, double* normfft_out) {
  if (!ctx || !out || !bc || !cf) return fail(CANSB200_EINVAL, "plan_create: null argument");
  for (int i = 0; i < 6; ++i)
    if (bc[i] != 'P' && bc[i] != 'D' && bc[i] != 'N') return fail(CANSB200_EINVAL, "plan_create: bc must be P, D or N");
  for (int d = 0; d < 3; ++d) {
    if (cf[d] != 'c' && cf[d] != 'f') return fail(CANSB200_EINVAL, "plan_create: c_or_f must be c or f");
    if ((bc[2 * d] == 'P') != (bc[2 * d + 1] == 'P')) return fail(CANSB200_EINVAL, "plan_create: periodic BCs must come in pairs");
  }
  auto pl = std::unique_ptr<cansb200_plan>(new cansb200_plan());
  pl->ctx = ctx;
  memcpy(pl->bc, bc, 6);
  memcpy(pl->cf, cf, 3);
  cansb200_options o;
  for (int* q = (int*)&o; q < (int*)(&o + 1); ++q) *q = -1;
  if (opt) o = *opt;
  pl->opt = o;
  // fftini, src/fft.f90:76-208
  double normfft = 1.0;
  for (int d = 0; d < 2; ++d) {
    double n1, n2;
    find_fft(bc[2 * d], bc[2 * d + 1], cf[d], pl->kind[d][0], pl->kind[d][1], n1, n2);
    const int ii = (bc[2 * d] == 'D' && bc[2 * d + 1] == 'D' && cf[d] == 'f') ? 1 : 0;
    pl->nt[d] = ctx->ng[d] - ii;
    if (pl->nt[d] < 1) return fail(CANSB200_EINVAL, "plan_create: transform length < 1");
    normfft *= n1 * (ctx->ng[d] + n2 - ii);
  }
  normfft = 1.0 / normfft;
  if (normfft_out) *normfft_out = ctx->is_fp32 ? (double)(1.0f / (float)(1.0 / normfft)) : normfft;
  // z solve geometry, src/solver.f90:77-82
  pl->q = (cf[2] == 'f' && bc[5] == 'D') ? 1 : 0;
  pl->periodic_z = (bc[4] == 'P' && bc[5] == 'P') ? 1 : 0;
  pl->th_n = ctx->ng[2] - pl->q;
  pl->th_nn = pl->periodic_z ? pl->th_n - 1 : pl->th_n;
  if (pl->th_nn < 1) return fail(CANSB200_EINVAL, "plan_create: z system is empty");
  // pipelined substitution: tiles of one 128-byte row segment (16 FP64 / 32 FP32 columns) x all rows, 1024 threads;
  // up to 512 rows one CTA per tile, up to 1024 rows a cluster of two CTAs shares the tile (variant 1) or one CTA
  // takes half as many columns (variant 2, and periodic z)
  int variant = o.thomas_variant >= 0 ? o.thomas_variant : 1;
  const int wide = ctx->is_fp32 ? 32 : 16;   // columns of a 128-byte row segment
  pl->th_cols = wide;
  pl->th_cl = 1;
  if (pl->th_nn > 512) {
    if (variant == 1 && !pl->periodic_z) pl->th_cl = 2;
    else pl->th_cols = wide / 2;
  }
  if (variant == 2) variant = 1;
  if (variant > 3) variant = 1;   // 3 = as 1, tiles fetched with cp.async instead of TMA
  const int chunks = CB_TH_THREADS / pl->th_cols * pl->th_cl;   // chunks per column over the whole cluster
  pl->th_m = (pl->th_nn + chunks - 1) / chunks;
  const int m_cap = ctx->is_fp32 ? 16 : 8;
  if (pl->th_m > m_cap) variant = 0;   // nz > 1024: the sequential kernel
  pl->th_variant = variant;
  if (!ctx->is_fp32) pl->th_mmax = (pl->th_cols == 16 && pl->th_cl == 1) ? (pl->th_m <= 4 ? 4 : 8) : (pl->th_m <= 6 ? 6 : 8);
  else pl->th_mmax = pl->th_cols == 16 ? 16 : (pl->th_cl == 1 ? (pl->th_m <= 8 ? 8 : 16) : (pl->th_m <= 12 ? 12 : 16));
  pl->nslots = o.cache_slots >= 1 ? (o.cache_slots > CB_MAX_SLOTS ? CB_MAX_SLOTS : o.cache_slots) : 1;
  {
    // shallow grids: tall tiles of jb y rows for the z-major one-GPU solve, so that a tile keeps ~512 rows
    const int chunks = CB_TH_THREADS / wide, nyz = ctx->n_z[1];
    if (ctx->nranks == 1 && (variant == 1 || variant == 3) && pl->th_cl == 1 && pl->th_cols == wide && !pl->periodic_z && pl->q == 0 &&
        pl->th_nn <= 256 && o.tall_tiles != 0) {
      int jb = 1;
      while (jb * 2 * pl->th_nn <= 512 && (nyz % (jb * 2)) == 0) jb *= 2;
      if (jb > 1) {
        pl->jb = jb;
        pl->th_m_b = (jb * pl->th_nn + chunks - 1) / chunks;
        if (!ctx->is_fp32) pl->th_mmax_b = pl->th_m_b <= 4 ? 4 : 8;
        else pl->th_mmax_b = pl->th_m_b <= 8 ? 8 : 16;
      }
    }
  }
  {
    // pivot-cache deduplication: a periodic direction whose length is a whole number of tile pairs (x), an even number of
    // rows held in full by this rank (y).  Off in distributed-TDMA mode (no pivot cache at all).
    const int nxz = ctx->n_z[0], nyz = ctx->n_z[1];
    // FP64 only: in FP32 the two halves of initsolver's eigenvalue array differ by up to 1e-3 relative (cancellation in
    // 1 - cos), far above the 1e-5 parity bar.  Not for the sequential variant, which is kept bit-identical to the reference.
    const bool on = o.pivot_dedup != 0 && !ctx->dtdma && variant != 0 && !ctx->is_fp32;
    int radix[4];
    const bool fastx = !ctx->force_generic && (ctx->is_fp32 ? r2r2_query<false, true>(nxz, 0, radix) : r2r2_query<false, false>(nxz, 0, radix)) > 0;
    pl->dx = on && fastx && bc[0] == 'P' && bc[1] == 'P' && nxz >= 4 * wide && (nxz % (2 * wide)) == 0;
    pl->dy = on && pl->jb == 1 && bc[2] == 'P' && bc[3] == 'P' && nyz == ctx->ng[1] && nyz >= 4 && (nyz % 2) == 0;   // (tall tiles need consecutive stored rows)
  }
  {
    const int rc = plan_cache_alloc(pl.get());
    if (rc) return rc;
  }
  // build the transform tables now so that the first solve does not allocate
  for (int d = 0; d < 2; ++d) {
    int rc;
    if (ctx->is_fp32) { FftTables<float>* t; rc = get_tables<float>(ctx, pl->nt[d], &t); }
    else { FftTables<double>* t; rc = get_tables<double>(ctx, pl->nt[d], &t); }
    if (rc) return rc;
  }
  ctx->nplans++;
  *out = pl.release();
  return 0;
}

int cansb200_plan_destroy(cansb200_plan* pl) {
  if (!pl) return 0;
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (auto& q : g_plan_ids)
      if (q == pl) q = nullptr;
  }
  pl->ctx->nplans--;
  pl->zcache.release(); pl->p2cache.release(); pl->dencache.release(); pl->state.release();
  pl->dtdma_big.release(); pl->dtdma_small.release(); pl->dtdma_rows.release();
  delete pl;
  return 0;
}

int cansb200_set_profiling(cansb200_ctx* c, int on) {
  if (!c) return fail(CANSB200_EINVAL, "null ctx");
  for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
  c->prof_events.clear();
  for (double& v : c->prof_ms) v = 0.0;
  c->prof_n = 0;
  c->profiling = on != 0;
  return 0;
}

int cansb200_get_profile(cansb200_ctx* c, double ms[8], unsigned long long* nsolves) {
  if (!c || !ms) return fail(CANSB200_EINVAL, "null argument");
  const size_t per = 7;
  CK(cudaDeviceSynchronize());
  const size_t ns = c->prof_events.size() / per;
  for (size_t s = 0; s < ns; ++s)
    for (size_t k = 0; k + 1 < per; ++k) {
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, c->prof_events[s * per + k], c->prof_events[s * per + k + 1]));
      c->prof_ms[k] += t;
    }
  c->prof_n += ns;
  for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
  c->prof_events.clear();
  for (int k = 0; k < 8; ++k) ms[k] = c->prof_ms[k];
  if (nsolves) *nsolves = c->prof_n;
  return 0;
}

int cansb200_plan_stats(cansb200_plan* pl, unsigned long long stats[4]) {
  if (!pl || !stats) return fail(CANSB200_EINVAL, "null argument");
  CacheState cs;
  CK(cudaMemcpy(&cs, pl->state.p, sizeof(cs), cudaMemcpyDeviceToHost));
  stats[0] = pl->solves; stats[1] = cs.nfactor; stats[2] = pl->ctx->launches; stats[3] = (unsigned long long)pl->th_variant | ((unsigned long long)pl->dx << 4) | ((unsigned long long)pl->dy << 5) |
             ((unsigned long long)(((pl->dx | pl->dy) && cs.sym_bad) ? 1 : 0) << 6) | ((unsigned long long)(pl->last_fused ? 1 : 0) << 7) |
             ((unsigned long long)pl->jb << 8);
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// device-side flags over the ranks of the box, stream ordered (dist_kernels.cuh): announce and / or wait for slot `slot`
// serialise the solves of one context on the device, whatever streams the caller uses
static int ctx_enter(cansb200_ctx* c, cudaStream_t st) {
  if (c->have_last && c->last_stream != st) CK(cudaStreamWaitEvent(st, c->last_done, 0));
  return 0;
}
static int ctx_leave(cansb200_ctx* c, cudaStream_t st) {
  if (!c->last_done) CK(cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming));
  CK(cudaEventRecord(c->last_done, st));
  c->last_stream = st;
  c->have_last = true;
  return 0;
}

static int dist_flag(cansb200_ctx* c, int slot, unsigned long long seq, bool signal, bool wait, cudaStream_t st, int target = -1,
                     bool skip_self = false) {
  DistPeers pp;
  for (int s = 0; s < CB_MAX_RANKS; ++s) pp.flags[s] = s < c->nranks ? (unsigned long long*)c->peer[s] : nullptr;
  dist_flag_kernel<<<1, 32, 0, st>>>(pp, c->rank, c->nranks, slot, seq, signal ? 1 : 0, wait ? 1 : 0, c->dist_status,
                                     c->dist_status_dev, 20ULL * 1000000000ULL, target, skip_self ? 1 : 0);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}
// a wait that timed out in an earlier solve left partly written buffers behind: refuse to go on
static int dist_check(cansb200_ctx* c, const char* who) {
  if (!c->connected) return fail(CANSB200_ECOMM, std::string(who) + ": cansb200_dist_connect has not been called");
  if (c->dist_status && *(volatile int*)c->dist_status)
    return fail(CANSB200_ECOMM, std::string(who) + ": a device-side wait for a peer timed out (a rank is missing or failed); results are invalid");
  return 0;
}

// can the y transforms of this plan store / load their rows through the peer row tables themselves (two-for-one kernels)?
template <class T> static bool dist_y_fast(cansb200_plan* pl) {
  cansb200_ctx* ctx = pl->ctx;
  if (ctx->force_generic || !kind_is_fast(pl->kind[1][0]) || !kind_is_fast(pl->kind[1][1])) return false;
  if ((ctx->ng[0] % (int)(16 / sizeof(T))) != 0) return false;   // column pairs move as 16-byte vectors
  int radix[4];
  return (sizeof(T) == 4 ? r2r2_query<true, true>(pl->nt[1], 0, radix) : r2r2_query<true, false>(pl->nt[1], 0, radix)) > 0;
}

// can the forward x transform of this plan take the fused fillps source?  (a two-for-one length over whole lines)
template <class T> static bool can_fuse_fillps(const cansb200_plan* pl) {
  const cansb200_ctx* ctx = pl->ctx;
  if (!ctx->fuse_fillps || ctx->force_generic || !kind_is_fast(pl->kind[0][0]) || pl->nt[0] != ctx->n[0]) return false;
  int radix[4];
  return r2r2_query<false, sizeof(T) == 4>(pl->nt[0], 0, radix) > 0;
}

// the z-slab decomposed solve: x and y transforms on my slab, tridiagonal stage on my z pencil.
//
// The two exchanges are the stores of the producing kernels (forward y transform -> z pencils of the owners,
// tridiagonal solve -> slabs of the owners).  The middle of the solve runs as a three-stage pipeline over W windows
// of x columns: while window w + 1 is still being transformed and sent (stage F, the caller's stream), window w is
// solved and sent back (stage T) and window w - 1 is transformed back (stage B).  The per-window flags replace the
// two whole-field barriers of round 1: NVLink (stages F and T) and HBM-only work (stage B, and the parts of F and T
// that read the local field) overlap instead of alternating.
template <class T>
static int solve_dist(cansb200_plan* pl, T* p, const int n[3], double normfft, const T* lam, const T* a, const T* b,
                      const T* c, bool y_packed, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  const void* ytab_fwd = y_packed ? ctx->ytab_fwd_pk.p : ctx->ytab_fwd.p;
  const void* ytab_bwd = y_packed ? ctx->ytab_bwd_pk.p : ctx->ytab_bwd.p;
  const unsigned long long seq = ++ctx->seq;   // first: every rank counts every collective call, whatever happens next
  int rc = dist_check(ctx, "solve");
  if (rc) return rc;
  const int nx = n[0], ny = n[1], nzl = n[2];
  const int nyl = ctx->n_z[1], nz = ctx->ng[2];
  const long long px = nx + 2, py = ny + 2;
  T* A = (T*)ctx->scratch.p;
  T* Cz = (T*)((char*)ctx->region + ctx->off_C);
  T* XB = (T*)((char*)ctx->region + ctx->off_XB);
  T* pin = p + (px * py + px + 1);
  if (pl->th_n != nz - pl->q) return fail(CANSB200_EINVAL, "solve: plan / grid mismatch");
  const bool pipelined = thomas_is_pipelined(pl);
  // Any y length / kind the single-GPU solve can transform also works here: when the two-for-one kernels do not serve it
  // (odd or prime n, REDFT00 / 11, RODFT00 / 11), the transform runs in place on the slab with the generic engine and a copy
  // kernel moves the rows through the same peer row tables (one extra pass; FFTW plans any n, src/fft.f90:148-162)
  const bool yfast = dist_y_fast<T>(pl);
  // windows: whole 128-byte row segments, as many as asked for (auto: 4) that divide nx
  const int wide = (int)(128 / sizeof(T));
  int W = 1;
  if (pipelined && !ctx->profiling && yfast) {
    // auto: measured on C3 (profiles/r2b_exchange_schedules.md) the windows pay on 2 ranks (2.82 -> 2.73 ms) and cost a
    // little on 4 and 8 (the kernels that store to the peers hold every SM while NVLink drains, whatever the schedule)
    const int want = ctx->dist_windows > 0 ? ctx->dist_windows : (ctx->nranks == 2 ? 4 : 1);
    for (W = want < CB_MAX_WINDOWS ? want : CB_MAX_WINDOWS; W > 1; --W)
      if (nx % (W * wide) == 0) break;
  }
  cudaStream_t sF = st, sT = st, sB = st;
  if (W > 1) {
    if (!ctx->dist_sT) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // later stages first: the pipeline drains instead of piling up
      CK(cudaStreamCreateWithPriority(&ctx->dist_sT, cudaStreamNonBlocking, hi));
      CK(cudaStreamCreateWithPriority(&ctx->dist_sB, cudaStreamNonBlocking, hi));
    }
    while ((int)ctx->dist_ev.size() < 2 * CB_MAX_WINDOWS + 1) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->dist_ev.push_back(e);
    }
    sT = ctx->dist_sT; sB = ctx->dist_sB;
  }
  // pivot-cache lookup (and factorisation on a miss) ahead of everything: it only needs the coefficients, and its one-time
  // read-back (first solve of a plan) must not sit behind a wait for the peers
  ThomasDev<T> D = make_thomas<T>(pl, nx, nyl, nx, (long long)nx * nyl, pl->th_n, pl->periodic_z, lam, a, b, c);
  D.out_rows = (const OutRow<T>*)ctx->ztab.p;
  rc = gaussel_prepare<T>(pl, D, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  R2RGeom gx{1, 1, px, nx, px * py, (long long)nx * ny, ny, nzl, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][0], pl->nt[0], pin, A, gx, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  const int ww = nx / W;
  // while the tridiagonal kernel shares the GPU with the y transforms of the neighbouring windows it must not take every
  // SM (one of its CTAs fills a whole SM): leave part of the machine to stages F and B
  const int cap = ctx->dist_thomas_ctas > 0 ? ctx->dist_thomas_ctas : (ctx->num_sms * 5) / 8;
  for (int w = 0; w < W; ++w) {
    const int xb = w * ww;
    // ---- stage F: forward y transform of the window; rows go straight to the z pencils of their owners
    R2RGeom gyf{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, ww, nzl, ny, 1};
    if (yfast) {
      gyf.row_tab = ytab_fwd;
      gyf.x0 = xb;
    }
    rc = run_r2r<T>(ctx, pl->kind[1][0], pl->nt[1], A + xb, A + xb, gyf, pl->opt.fft_y_lines, sF);
    if (rc) return rc;
    if (!yfast) {
      slab_rows_copy_kernel<T><<<(unsigned)(ctx->num_sms * 8), 256, 0, sF>>>(A, nx, (long long)nx * ny, (const DistRow<T>*)ytab_fwd, nx, ny, nzl, 1);
      ctx->launches++;
      CK(cudaGetLastError());
    }
    if (W > 1) {
      rc = dist_flag(ctx, CB_SLOT_FWD + w, seq, true, false, sF);
      if (rc) return rc;
      CK(cudaEventRecord(ctx->dist_ev[w], sF));
      CK(cudaStreamWaitEvent(sT, ctx->dist_ev[w], 0));
      rc = dist_flag(ctx, CB_SLOT_FWD + w, seq, false, true, sT);
    } else {
      rc = dist_flag(ctx, CB_SLOT_FWD, seq, true, true, sT);
    }
    if (rc) return rc;
    // ---- stage T: tridiagonal solve on the window of my z pencil (nx, ny/P, nz); result rows go to the slabs
    if (W == 1) {
      prof_mark(ctx, st);
      prof_mark(ctx, st);
    }
    if (pipelined) {
      D.xb = xb; D.xn = ww;
      ctx->cta_cap = W > 1 ? cap : 0;
      rc = gaussel_apply<T>(pl, D, Cz, (T)normfft, sT);
      ctx->cta_cap = 0;
      if (rc) return rc;
      if (pl->th_n < nz) {   // face-centred Dirichlet: the last plane is not part of the system but still travels back
        scatter_rows_kernel<T><<<ctx->num_sms, 256, 0, sT>>>(Cz, D.sk, (const DistOutRow<T>*)D.out_rows, pl->th_n, nz, nyl, nx, xb, ww);
        ctx->launches++;
      }
    } else {
      ThomasDev<T> D2 = D;
      D2.out_rows = nullptr;
      rc = gaussel_apply<T>(pl, D2, Cz, (T)normfft, sT);
      if (rc) return rc;
      scatter_rows_kernel<T><<<ctx->num_sms * 4, 256, 0, sT>>>(Cz, D.sk, (const DistOutRow<T>*)D.out_rows, 0, nz, nyl, nx, 0, nx);
      ctx->launches++;
    }
    CK(cudaGetLastError());
    if (W > 1) {
      rc = dist_flag(ctx, CB_SLOT_BWD + w, seq, true, false, sT);
      if (rc) return rc;
      CK(cudaEventRecord(ctx->dist_ev[CB_MAX_WINDOWS + w], sT));
      CK(cudaStreamWaitEvent(sB, ctx->dist_ev[CB_MAX_WINDOWS + w], 0));
      rc = dist_flag(ctx, CB_SLOT_BWD + w, seq, false, true, sB);
    } else {
      rc = dist_flag(ctx, CB_SLOT_BWD, seq, true, true, sB);
    }
    if (rc) return rc;
    if (W == 1) prof_mark(ctx, st);
    // ---- stage B: backward y transform of the window; gathers its rows from the way-back buffer
    R2RGeom gyb{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, ww, nzl, ny, 1};
    if (yfast) {
      gyb.row_tab = ytab_bwd;
      gyb.x0 = xb;
      rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], XB, A + xb, gyb, pl->opt.fft_y_lines, sB);
    } else {
      slab_rows_copy_kernel<T><<<(unsigned)(ctx->num_sms * 8), 256, 0, sB>>>(A, nx, (long long)nx * ny, (const DistRow<T>*)ytab_bwd, nx, ny, nzl, 0);
      ctx->launches++;
      CK(cudaGetLastError());
      rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], A, A, gyb, pl->opt.fft_y_lines, sB);
    }
    if (rc) return rc;
  }
  if (W > 1) {
    CK(cudaEventRecord(ctx->dist_ev[2 * CB_MAX_WINDOWS], sB));
    CK(cudaStreamWaitEvent(st, ctx->dist_ev[2 * CB_MAX_WINDOWS], 0));
  }
  prof_mark(ctx, st);
  R2RGeom gxb{1, 1, nx, px, (long long)nx * ny, px * py, ny, nzl, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][1], pl->nt[0], A, pin, gxb, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  return 0;
}

// ---------------------------------------------------------------------------
// Copy-engine flavour of the exchange (CANSB200_CTX_DIST_MODE = 1).
//
// Measured on this box (profiles/r2a_nvlink_store_patterns.txt): SM stores fill NVLink to ~690 GB/s per direction in any
// pattern, but only with ~300 resident CTAs -- a producing kernel that stores to its peers holds the whole GPU while the
// link drains, so the HBM-only stages (x transforms, backward y transform) cannot overlap it.  A copy-engine transfer
// moves 700-760 GB/s (dense blocks, or 2-D blocks with >= 2 KB rows) with no SM involved.  So here the producers write
// dense per-destination blocks into LOCAL send buffers (same kernels, the row tables simply point at local memory; the
// block for this rank itself goes straight to its final place) and one copy stream per peer moves them:
//   forward  : z chunks.  x transform + y transform of chunk c + 1 run while chunk c is on the wire; a chunk's block for
//              peer s is contiguous on both sides ([plane][row][x]).
//   backward : x windows.  Tridiagonal solve of window w + 1, transfer of window w (2-D copy, rows of window width) and
//              backward y transform of window w - 1 run concurrently.
// Every transfer is announced to its receiver by a flag written on the copy stream behind it.
template <class T> static int build_dma_tables(cansb200_ctx* c) {
  const int P = c->nranks, r = c->rank;
  const long long nx = c->ng[0];
  const int ny = c->ng[1], nz = c->ng[2];
  const long long nzl_r = c->zs[r + 1] - c->zs[r], nyl_r = c->ys[r + 1] - c->ys[r];
  if (c->scratch2.ensure((size_t)nx * ny * nzl_r * sizeof(T)) || c->sendb.ensure((size_t)nx * nyl_r * nz * sizeof(T)))
    return fail(CANSB200_ENOMEM, "solve: send buffers of the copy-engine exchange");
  T* SF = (T*)c->scratch2.p;
  T* SB = (T*)c->sendb.p;
  T* C_r = (T*)((char*)c->region + c->off_C);
  T* XB_r = (T*)((char*)c->region + c->off_XB);
  for (int packed = 0; packed < 2; ++packed) {
    std::vector<R2Row<T>> yf(ny);
    for (int j = 0; j < ny; ++j) {
      const int pj = packed ? pack_index(j, ny) : j;
      int s = 0;
      while (pj >= c->ys[s + 1]) ++s;
      const long long nyl_s = c->ys[s + 1] - c->ys[s], jl = pj - c->ys[s];
      // forward y output row j of my plane g: block s of the send buffer, [g][jl][i]; my own block goes home directly
      yf[j].ptr = s == r ? C_r + ((long long)c->zs[r] * nyl_s + jl) * nx : SF + nx * nzl_r * c->ys[s] + jl * nx;
      yf[j].gs = nyl_s * nx;
    }
    DevBuf& tf = packed ? c->ytab_fwd_loc_pk : c->ytab_fwd_loc;
    if (tf.ensure(sizeof(R2Row<T>) * ny)) return fail(CANSB200_ENOMEM, "solve: tables");
    CK(upload(tf.p, yf.data(), sizeof(R2Row<T>) * ny));
  }
  std::vector<OutRow<T>> zt(nz);
  for (int s = 0; s < P; ++s) {
    const long long nzl_s = c->zs[s + 1] - c->zs[s];
    for (int k = c->zs[s]; k < c->zs[s + 1]; ++k) {   // result row k of my columns: block s of the send buffer, [jl][k - zs[s]][i]
      zt[k].ptr = s == r ? XB_r + nzl_r * c->ys[r] * nx + (long long)(k - c->zs[s]) * nx
                         : SB + nx * nyl_r * c->zs[s] + (long long)(k - c->zs[s]) * nx;
      zt[k].sj = nzl_s * nx;
    }
  }
  if (c->ztab_loc.ensure(sizeof(OutRow<T>) * nz)) return fail(CANSB200_ENOMEM, "solve: tables");
  CK(upload(c->ztab_loc.p, zt.data(), sizeof(OutRow<T>) * nz));
  int lo = 0, hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  c->dist_cs.assign(P, nullptr);
  for (int s = 0; s < P; ++s)
    if (s != r) CK(cudaStreamCreateWithPriority(&c->dist_cs[s], cudaStreamNonBlocking, hi));
  return 0;
}

// Two-half schedule of the distributed solve.  dma = true: the copy-engine exchange described above.  dma = false: the
// producing kernels store straight into the peers' buffers (as solve_dist does), but in the same two halves -- z chunks
// forward, so that the x transform of chunk c + 1 (HBM only) runs next to the y transform of chunk c (NVLink bound; its
// CTAs are padded to one per SM so that it leaves room), x windows backward (tridiagonal kernel on part of the SMs next to
// the backward y transform of the previous window).
template <class T>
static int solve_dist2(cansb200_plan* pl, T* p, const int n[3], double normfft, const T* lam, const T* a, const T* b,
                       const T* c, bool y_packed, bool dma, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  const unsigned long long seq = ++ctx->seq;
  int rc = dist_check(ctx, "solve");
  if (rc) return rc;
  const int P = ctx->nranks, r = ctx->rank;
  const int nx = n[0], ny = n[1], nzl = n[2];
  const int nyl = ctx->n_z[1], nz = ctx->ng[2];
  const long long px = nx + 2, py = ny + 2;
  if (pl->th_n != nz - pl->q) return fail(CANSB200_EINVAL, "solve: plan / grid mismatch");
  if (dma && ctx->dist_cs.empty()) {
    rc = build_dma_tables<T>(ctx);
    if (rc) return rc;
  }
  T* A = (T*)ctx->scratch.p;
  T* SF = (T*)ctx->scratch2.p;
  T* SB = (T*)ctx->sendb.p;
  T* Cz = (T*)((char*)ctx->region + ctx->off_C);
  T* XB = (T*)((char*)ctx->region + ctx->off_XB);
  T* pin = p + (px * py + px + 1);
  const void* ytab_fwd = dma ? (y_packed ? ctx->ytab_fwd_loc_pk.p : ctx->ytab_fwd_loc.p) : (y_packed ? ctx->ytab_fwd_pk.p : ctx->ytab_fwd.p);
  const void* ytab_bwd = y_packed ? ctx->ytab_bwd_pk.p : ctx->ytab_bwd.p;
  const int wide = (int)(128 / sizeof(T));
  // z chunks of the forward half, x windows of the backward half
  int Cn = ctx->profiling ? 1 : (ctx->dist_chunks > 0 ? ctx->dist_chunks : 4);
  if (Cn > nzl) Cn = nzl;
  int W = 1;
  if (!ctx->profiling) {
    const int want = ctx->dist_windows > 0 ? ctx->dist_windows : 4;
    for (W = want < CB_MAX_WINDOWS ? want : CB_MAX_WINDOWS; W > 1; --W)
      if (nx % (W * wide) == 0) break;
  }
  const bool pipelined = thomas_is_pipelined(pl);
  if (!pipelined) W = 1;   // the sequential kernel has no column windows
  if (!ctx->dist_sT) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ctx->dist_sT, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&ctx->dist_sB, cudaStreamNonBlocking, hi));
  }
  if (!ctx->dist_sF) CK(cudaStreamCreateWithFlags(&ctx->dist_sF, cudaStreamNonBlocking));
  while ((int)ctx->dist_ev.size() < 2 * CB_MAX_WINDOWS + 2) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->dist_ev.push_back(e);
  }
  cudaStream_t sT = ctx->dist_sT, sB = ctx->dist_sB;
  cudaStream_t sF = (dma || Cn == 1) ? st : ctx->dist_sF;   // stream of the forward y transforms
  const size_t esz = sizeof(T);
  const int pad_kb = dma || Cn == 1 ? 0 : (ctx->dist_split_pad >= 0 ? ctx->dist_split_pad : 52);
  prof_mark(ctx, st);
  ThomasDev<T> D = make_thomas<T>(pl, nx, nyl, nx, (long long)nx * nyl, pl->th_n, pl->periodic_z, lam, a, b, c);
  D.out_rows = (const OutRow<T>*)(dma ? ctx->ztab_loc.p : ctx->ztab.p);
  rc = gaussel_prepare<T>(pl, D, st);   // pivot-cache lookup (factorisation on a miss) ahead of everything
  if (rc) return rc;
  // ---- forward half: x transform, y transform (into the send blocks / the peers' z pencils), chunk by chunk
  for (int q = 0; q < Cn; ++q) {
    const int g0 = (int)((long long)nzl * q / Cn), g1 = (int)((long long)nzl * (q + 1) / Cn), ng_ = g1 - g0;
    if (ng_ < 1) continue;
    R2RGeom gx{1, 1, px, nx, px * py, (long long)nx * ny, ny, ng_, nx, 0};
    rc = run_r2r<T>(ctx, pl->kind[0][0], pl->nt[0], pin + (long long)g0 * px * py, A + (long long)g0 * nx * ny, gx, pl->opt.fft_x_lines, st);
    if (rc) return rc;
    if (Cn == 1) prof_mark(ctx, st);
    if (sF != st) {
      CK(cudaEventRecord(ctx->dist_ev[q], st));
      CK(cudaStreamWaitEvent(sF, ctx->dist_ev[q], 0));
    }
    R2RGeom gyf{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, nx, ng_, ny, 1};
    gyf.row_tab = ytab_fwd;
    gyf.g0 = g0;
    gyf.smem_pad = pad_kb * 1024;
    rc = run_r2r<T>(ctx, pl->kind[1][0], pl->nt[1], A + (long long)g0 * nx * ny, A + (long long)g0 * nx * ny, gyf, pl->opt.fft_y_lines, sF);
    if (rc) return rc;
    if (!dma) continue;
    CK(cudaEventRecord(ctx->dist_ev[q], st));
    for (int d = 1; d < P; ++d) {
      const int s = (r + d) % P;   // every rank starts with a different peer
      const long long nyl_s = ctx->ys[s + 1] - ctx->ys[s];
      T* dst = (T*)((char*)ctx->peer[s] + ctx->off_C) + ((long long)(ctx->zs[r] + g0) * nyl_s) * nx;
      const T* src = SF + (long long)nx * nzl * ctx->ys[s] + (long long)g0 * nyl_s * nx;
      CK(cudaStreamWaitEvent(ctx->dist_cs[s], ctx->dist_ev[q], 0));
      CK(cudaMemcpyAsync(dst, src, (size_t)ng_ * nyl_s * nx * esz, cudaMemcpyDeviceToDevice, ctx->dist_cs[s]));
    }
  }
  if (dma) {
    for (int d = 1; d < P; ++d) {
      const int s = (r + d) % P;
      rc = dist_flag(ctx, CB_SLOT_FWD, seq, true, false, ctx->dist_cs[s], s);   // "my block has landed on your z pencil"
      if (rc) return rc;
    }
    CK(cudaStreamWaitEvent(sT, ctx->dist_ev[Cn - 1], 0));   // my own block is written by my own kernels
    rc = dist_flag(ctx, CB_SLOT_FWD, seq, false, true, sT, -1, true);
  } else {
    rc = dist_flag(ctx, CB_SLOT_FWD, seq, true, false, sF);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->dist_ev[2 * CB_MAX_WINDOWS + 1], sF));
    CK(cudaStreamWaitEvent(sT, ctx->dist_ev[2 * CB_MAX_WINDOWS + 1], 0));
    rc = dist_flag(ctx, CB_SLOT_FWD, seq, false, true, sT);
  }
  if (rc) return rc;
  prof_mark(ctx, sT);
  prof_mark(ctx, sT);
  // ---- backward half: tridiagonal solve of an x window (into the send blocks / the peers' slabs), backward y transform
  const int ww = nx / W;
  const int cap = ctx->dist_thomas_ctas > 0 ? ctx->dist_thomas_ctas : (ctx->num_sms * 5) / 8;
  for (int w = 0; w < W; ++w) {
    const int xb = w * ww;
    if (pipelined) {
      D.xb = xb; D.xn = ww;
      ctx->cta_cap = (!dma && W > 1) ? cap : 0;
      rc = gaussel_apply<T>(pl, D, Cz, (T)normfft, sT);
      ctx->cta_cap = 0;
      if (rc) return rc;
      if (pl->th_n < nz) {
        scatter_rows_kernel<T><<<ctx->num_sms, 256, 0, sT>>>(Cz, D.sk, (const DistOutRow<T>*)D.out_rows, pl->th_n, nz, nyl, nx, xb, ww);
        ctx->launches++;
      }
    } else {
      ThomasDev<T> D2 = D;
      D2.out_rows = nullptr;
      rc = gaussel_apply<T>(pl, D2, Cz, (T)normfft, sT);
      if (rc) return rc;
      scatter_rows_kernel<T><<<ctx->num_sms * 4, 256, 0, sT>>>(Cz, D.sk, (const DistOutRow<T>*)D.out_rows, 0, nz, nyl, nx, 0, nx);
      ctx->launches++;
    }
    CK(cudaGetLastError());
    if (!dma) {
      rc = dist_flag(ctx, CB_SLOT_BWD + w, seq, true, false, sT);
      if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->dist_ev[CB_MAX_WINDOWS + w], sT));
    for (int d = 1; dma && d < P; ++d) {
      const int s = (r + d) % P;
      const long long nzl_s = ctx->zs[s + 1] - ctx->zs[s];
      T* dst = (T*)((char*)ctx->peer[s] + ctx->peer_off_XB[s]) + nzl_s * ctx->ys[r] * nx + xb;
      const T* src = SB + (long long)nx * nyl * ctx->zs[s] + xb;
      CK(cudaStreamWaitEvent(ctx->dist_cs[s], ctx->dist_ev[CB_MAX_WINDOWS + w], 0));
      if (W == 1)
        CK(cudaMemcpyAsync(dst, src, (size_t)nyl * nzl_s * nx * esz, cudaMemcpyDeviceToDevice, ctx->dist_cs[s]));
      else
        CK(cudaMemcpy2DAsync(dst, (size_t)nx * esz, src, (size_t)nx * esz, (size_t)ww * esz, (size_t)nyl * nzl_s, cudaMemcpyDeviceToDevice,
                             ctx->dist_cs[s]));
      rc = dist_flag(ctx, CB_SLOT_BWD + w, seq, true, false, ctx->dist_cs[s], s);
      if (rc) return rc;
    }
    CK(cudaStreamWaitEvent(sB, ctx->dist_ev[CB_MAX_WINDOWS + w], 0));
    rc = dist_flag(ctx, CB_SLOT_BWD + w, seq, false, true, sB, -1, dma);
    if (rc) return rc;
    if (W == 1) prof_mark(ctx, sB);
    R2RGeom gyb{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, ww, nzl, ny, 1};
    gyb.row_tab = ytab_bwd;
    gyb.x0 = xb;
    rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], XB, A + xb, gyb, pl->opt.fft_y_lines, sB);
    if (rc) return rc;
  }
  CK(cudaEventRecord(ctx->dist_ev[2 * CB_MAX_WINDOWS], sB));
  CK(cudaStreamWaitEvent(st, ctx->dist_ev[2 * CB_MAX_WINDOWS], 0));
  prof_mark(ctx, st);
  R2RGeom gxb{1, 1, nx, px, (long long)nx * ny, px * py, ny, nzl, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][1], pl->nt[0], A, pin, gxb, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  return 0;
}

// the z-slab decomposed solve with the distributed TDMA (is_poisson_dtdma): no transposes; the only exchange is the
// 2-rows-per-rank reduced system, gathered on every rank (peer stores into the otherwise unused z-pencil buffer) and
// solved redundantly.  lam = lambdaxy(nx, ny) of the whole slab, a / b / c = my z slice (cansb200_ctx_set DTDMA).
template <class T>
static int solve_dist_dtdma(cansb200_plan* pl, T* p, const int n[3], double normfft, const T* lam, const T* a, const T* b,
                            const T* c, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  const unsigned long long seq = ++ctx->seq;
  int rc = dist_check(ctx, "solve (dtdma)");
  if (rc) return rc;
  const int nx = n[0], ny = n[1], nzl = n[2], P = ctx->nranks, r = ctx->rank;
  const long long px = nx + 2, py = ny + 2;
  const size_t ncol = (size_t)nx * ny;
  const int nloc = nzl - (r == P - 1 ? pl->q : 0);
  if (nloc < 3) return fail(CANSB200_EUNSUPPORTED, "solve (dtdma): every rank needs at least 3 z planes");
  const size_t nel_z = (size_t)ctx->ng[0] * (ctx->ys[r + 1] - ctx->ys[r]) * ctx->ng[2];
  // the gather buffer [3][2 P][ncol] lives in the z-pencil slot of the exchange region (same offset on every rank);
  // the smallest slot of the box must hold it
  size_t min_slot = nel_z;
  for (int s = 0; s < P; ++s) {
    const size_t e = (size_t)ctx->ng[0] * (ctx->ys[s + 1] - ctx->ys[s]) * ctx->ng[2];
    if (e < min_slot) min_slot = e;
  }
  if ((size_t)6 * P * ncol > min_slot) return fail(CANSB200_EUNSUPPORTED, "solve (dtdma): nz too small for the reduced-system buffer (needs nz >= 6 P^2)");
  // coefficient cache: Z / AA / CC and the reduced coefficient rows per slot, keyed by the content of (a, b, c, lambda)
  const int nslots = pl->nslots;
  const size_t slot_big = 3 * ncol * (size_t)nloc, slot_small = 5 * ncol, slot_rows = 2 * (size_t)P * ncol;
  if (pl->dtdma_big.ensure(slot_big * nslots * sizeof(T))) return fail(CANSB200_ENOMEM, "solve (dtdma): coefficient arrays");
  if (pl->dtdma_small.ensure((slot_small * nslots + (2 + 4 * (size_t)P) * ncol) * sizeof(T))) return fail(CANSB200_ENOMEM, "solve (dtdma): reduced system");
  if (pl->dtdma_rows.ensure(2 * slot_rows * nslots * sizeof(T))) return fail(CANSB200_ENOMEM, "solve (dtdma): gathered coefficient rows");
  T* A = (T*)ctx->scratch.p;
  T* G = (T*)((char*)ctx->region + ctx->off_C);
  T* pin = p + (px * py + px + 1);
  const CacheState* cst = (const CacheState*)pl->state.p;
  {
    // content hash -> slot (the same device-side machinery as the pivot cache of the transposed solve)
    ThomasDev<T> H = make_thomas<T>(pl, nx, ny, nx, (long long)nx * ny, nloc, 0, lam, a, b, c);
    H.dx = H.dy = 0; H.nxu = nx; H.nyu = ny;
    const long long total = 3LL * nloc + (long long)ncol;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 592) blocks = 592;
    thomas_hash_kernel<T><<<blocks, 256, 0, st>>>(H, (CacheState*)pl->state.p);
    thomas_select_kernel<<<1, 32, 0, st>>>((CacheState*)pl->state.p);
    ctx->launches += 2;
  }
  prof_mark(ctx, st);
  R2RGeom gx{1, 1, px, nx, px * py, (long long)nx * ny, ny, nzl, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][0], pl->nt[0], pin, A, gx, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  R2RGeom gy{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, nx, nzl, ny, 1};
  rc = run_r2r<T>(ctx, pl->kind[1][0], pl->nt[1], A, A, gy, pl->opt.fft_y_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  // local elimination on my slab
  DtdmaDev<T> D;
  D.nx = nx; D.ny = ny; D.n = nloc; D.nranks = 1; D.periodic = 0;
  D.starts[0] = 0; D.starts[1] = nloc;
  D.a = a; D.b = b; D.c = c; D.lam = lam;
  D.st = cst; D.slot_big = (long long)slot_big; D.slot_small = (long long)slot_small;
  T* big = (T*)pl->dtdma_big.p;
  D.Z = big; D.AA = big + ncol * nloc; D.CC = big + 2 * ncol * nloc;
  T* sm = (T*)pl->dtdma_small.p;
  D.Z1 = sm; D.ra = sm + ncol; D.rc = sm + 3 * ncol;
  T* tail = sm + slot_small * nslots;
  D.rp = tail; D.rcw = tail + 2 * ncol; D.rp2 = D.rcw + 2 * ncol * P;
  const unsigned cbk = (unsigned)((ncol + 127) / 128);
  dtdma_coef_kernel<T><<<cbk, 128, 0, st>>>(D);   // returns at once on a cache hit
  prof_mark(ctx, st);
  {
    // slab-local elimination: on chip with the pipelined tridiagonal kernel (tile = one 128-byte row segment of columns x all
    // rows of the slab: p in, Z in, p out = 24 B/point) whenever the slab fits its chunking, else the one-thread-per-column
    // sweeps through HBM (48 B/point)
    const int wide = (int)(128 / sizeof(T)), chunks = CB_TH_THREADS / wide;
    const int m = (nloc + chunks - 1) / chunks;
    const int m_lo = ctx->is_fp32 ? 8 : 4, m_hi = ctx->is_fp32 ? 16 : 8;
    // ... when it pays: a tile is one 128-byte row segment x the slab's rows, and below ~4 rows per thread (slabs under 256
    // rows in FP64) the per-tile overhead of the chunked scan exceeds the traffic it saves (measured on 8 B200: C3 with 64-row
    // slabs 1.57 ms per solve tiled against 1.1 ms with the per-column sweeps; C5 with 128-row slabs 9.7 against 8.4 ms)
    const bool pays = m >= m_lo || ctx->dtdma_tiled == 2;
    const bool tiled = ctx->dtdma_tiled && pays && m <= m_hi && (nx % wide) == 0 && nx >= wide;
    if (tiled) {
      ThomasDev<T> Dt = make_thomas<T>(pl, nx, ny, nx, (long long)ncol, nloc, 0, lam, a, b, c);
      Dt.m = m; Dt.nopin = 1; Dt.dx = Dt.dy = 0; Dt.nxu = nx; Dt.nyu = ny;
      Dt.zsj = nx; Dt.zsk = (long long)ncol;
      Dt.dt_mode = 1; Dt.dt_z1 = D.Z1; Dt.dt_rp = D.rp; Dt.dt_slot_small = (long long)slot_small;
      // borrow the launcher of the transposed solve with this slab's chunking
      const int sv_m = pl->th_m, sv_mmax = pl->th_mmax, sv_cols = pl->th_cols, sv_cl = pl->th_cl, sv_var = pl->th_variant;
      pl->th_m = m; pl->th_mmax = m <= m_lo ? m_lo : m_hi; pl->th_cols = wide; pl->th_cl = 1; pl->th_variant = 1;
      pl->zsrc_base = big; pl->zsrc_slot = (long long)slot_big;
      rc = gaussel_apply<T>(pl, Dt, A, (T)normfft, st);
      pl->zsrc_base = nullptr; pl->zsrc_slot = 0;
      pl->th_m = sv_m; pl->th_mmax = sv_mmax; pl->th_cols = sv_cols; pl->th_cl = sv_cl; pl->th_variant = sv_var;
      if (rc) return rc;
    } else {
      dtdma_phase1_kernel<T><<<cbk, 128, 0, st>>>(D, A, (T)normfft);
      ctx->launches++;
    }
  }
  DtdmaPeers pp;
  for (int s = 0; s < CB_DTDMA_MAX_RANKS; ++s) pp.dst[s] = s < P ? (void*)((char*)ctx->peer[s] + ctx->off_C) : nullptr;
  dtdma_gather_kernel<T><<<ctx->num_sms * 4, 256, 0, st>>>(pp, r, P, (long long)ncol, D.ra, D.rc, D.rp, cst, (long long)slot_small);
  ctx->launches += 2;
  CK(cudaGetLastError());
  rc = dist_flag(ctx, CB_SLOT_BAR, seq, true, true, st);
  if (rc) return rc;
  // reduced system of all ranks, solved redundantly on every rank; then my inner rows
  T* GA = (T*)pl->dtdma_rows.p;
  T* GC = GA + slot_rows * nslots;
  dtdma_save_rows_kernel<T><<<ctx->num_sms * 2, 256, 0, st>>>(G, (long long)slot_rows, GA, GC, cst, (long long)slot_rows);
  DtdmaDev<T> R = D;
  R.nranks = P; R.periodic = pl->periodic_z;
  R.ra = GA; R.rc = GC; R.slot_small = (long long)slot_rows;
  R.rp = G + 4 * (size_t)P * ncol;
  dtdma_reduced_kernel<T><<<cbk, 128, 0, st>>>(R);
  DtdmaDev<T> F = D;
  F.rp = R.rp + 2 * (size_t)r * ncol;
  dtdma_phase3_kernel<T><<<ctx->num_sms * 8, 256, 0, st>>>(F, A);
  ctx->launches += 3;
  CK(cudaGetLastError());
  rc = dist_flag(ctx, CB_SLOT_BAR + 1, seq, true, true, st);   // nobody refills my gather buffer before I have read it
  if (rc) return rc;
  prof_mark(ctx, st);
  rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], A, A, gy, pl->opt.fft_y_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  R2RGeom gxb{1, 1, nx, px, (long long)nx * ny, px * py, ny, nzl, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][1], pl->nt[0], A, pin, gxb, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  return 0;
}

// ---------------------------------------------------------------------------
template <class T>
static int solve_impl(cansb200_plan* pl, void* p_any, const int n[3], double normfft, const void* lam_any, const void* a_any,
                      const void* b_any, const void* c_any, int mem_kind, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  NvtxRange nvtx_solve("cansb200_solve");
  const int nx = n[0], ny = n[1], nz = n[2];
  const long long px = nx + 2, py = ny + 2;          // haloed pitches
  const size_t nh = (size_t)px * py * (nz + 2);
  T* p = (T*)p_any;
  const T *lam = (const T*)lam_any, *a = (const T*)a_any, *b = (const T*)b_any, *c = (const T*)c_any;
  // host-memory mode, one rank: chunked copies overlapped with the transforms (see cansb200_ctx::host_chunks)
  const bool chunked = mem_kind == CANSB200_MEM_HOST && ctx->nranks == 1 && ctx->host_chunks > 1 && nz >= 2 && !ctx->profiling;
  const size_t plane = (size_t)px * py;
  if (mem_kind == CANSB200_MEM_HOST && ctx->pin_host) {
    // a pageable Fortran array copies at a fraction of the PCIe rate: page-lock it once (CaNS allocates p once per run)
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p_any) == cudaSuccess && at.type == cudaMemoryTypeUnregistered) {
      if (cudaHostRegister(p_any, nh * sizeof(T), cudaHostRegisterDefault) == cudaSuccess)
        ctx->pinned.emplace_back(p_any, nh * sizeof(T));
      else
        cudaGetLastError();   // not fatal: the copies just stay pageable
    } else {
      cudaGetLastError();
    }
  }
  if (mem_kind == CANSB200_MEM_HOST) {
    if (ctx->staging.ensure(nh * sizeof(T))) return fail(CANSB200_ENOMEM, "solve: staging");
    const size_t nzg = (size_t)ctx->n_z[2], nlam = (size_t)ctx->n_z[0] * ctx->n_z[1];   // a, b, c(n_z(3)); lambdaxy(n_z(1), n_z(2))
    const size_t ncoef = 3 * nzg + nlam;
    if (ctx->coef.ensure(ncoef * sizeof(T))) return fail(CANSB200_ENOMEM, "solve: coefficient staging");
    T* cf = (T*)ctx->coef.p;
    if (!chunked) CK(cudaMemcpyAsync(ctx->staging.p, p_any, nh * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf, a_any, nzg * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf + nzg, b_any, nzg * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf + 2 * nzg, c_any, nzg * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf + 3 * nzg, lam_any, nlam * sizeof(T), cudaMemcpyHostToDevice, st));
    p = (T*)ctx->staging.p;
    a = cf; b = cf + nzg; c = cf + 2 * nzg; lam = cf + 3 * nzg;
  }
  // eigenvalues in the order of an _OPENACC-built initsolver (option lambda_order = 1): bring them to halfcomplex order.
  // Several ranks: x only -- the y rows are dealt out in packed order instead (build_dist_tables), so that the local
  // slice of lambdaxy means what it means in the reference.
  struct XsplitGuard {
    cansb200_ctx* c;
    XsplitGuard(cansb200_ctx* c_, int v) : c(c_) { c->cur_xsplit = v; }
    ~XsplitGuard() { c->cur_xsplit = 0; }
  } xsplit_guard(ctx, pl->dx);
  const bool pk = pl->opt.lambda_order == 1;
  const bool pkx = pk && pl->bc[0] == 'P', pky = pk && pl->bc[2] == 'P';
  const bool pky_local = pky && (ctx->nranks == 1 || ctx->dtdma);
  if (pkx || pky_local || pl->dx) {
    const int lx = ctx->n_z[0], ly = ctx->n_z[1];
    if (ctx->lam_perm.ensure((size_t)lx * ly * sizeof(T))) return fail(CANSB200_ENOMEM, "solve: eigenvalue buffer");
    const long long tot = (long long)lx * ly;
    lambda_unpack_kernel<T><<<(unsigned)((tot + 255) / 256 < 1184 ? (tot + 255) / 256 : 1184), 256, 0, st>>>(lam, (T*)ctx->lam_perm.p, lx, ly,
                                                                                                          pkx ? 1 : 0, pky_local ? 1 : 0, pl->dx);
    ctx->launches++;
    CK(cudaGetLastError());
    lam = (const T*)ctx->lam_perm.p;
  }
  if (chunked) {
    const int nch = ctx->host_chunks < nz ? ctx->host_chunks : nz;
    if (!ctx->h2d_stream) CK(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    while ((int)ctx->chunk_ev.size() < 2 * nch + 1) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->chunk_ev.push_back(e);
    }
    T* A = (T*)ctx->scratch.p;
    T* hp = (T*)p_any;
    // the copy stream starts after whatever the caller's stream has queued
    CK(cudaEventRecord(ctx->chunk_ev[2 * nch], st));
    CK(cudaStreamWaitEvent(ctx->h2d_stream, ctx->chunk_ev[2 * nch], 0));
    int rc;
    auto k_lo = [&](int q) { return (int)((long long)nz * q / nch); };
    for (int q = 0; q < nch; ++q) {
      const int k0 = k_lo(q), k1 = k_lo(q + 1), nk = k1 - k0;
      // interior planes k0+1 .. k1 of the haloed array (whole xy planes, x / y halos included)
      CK(cudaMemcpyAsync(p + (size_t)(k0 + 1) * plane, hp + (size_t)(k0 + 1) * plane, (size_t)nk * plane * sizeof(T),
                         cudaMemcpyHostToDevice, ctx->h2d_stream));
      CK(cudaEventRecord(ctx->chunk_ev[q], ctx->h2d_stream));
      CK(cudaStreamWaitEvent(st, ctx->chunk_ev[q], 0));
      T* pin = p + ((size_t)(k0 + 1) * plane + px + 1);
      T* Aq = A + (size_t)k0 * nx * ny;
      R2RGeom gx{1, 1, px, nx, px * py, (long long)nx * ny, ny, nk, nx, 0};
      rc = run_r2r<T>(ctx, pl->kind[0][0], pl->nt[0], pin, Aq, gx, pl->opt.fft_x_lines, st);
      if (rc) return rc;
      R2RGeom gy{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, nx, nk, ny, 1};
      rc = run_r2r<T>(ctx, pl->kind[1][0], pl->nt[1], Aq, Aq, gy, pl->opt.fft_y_lines, st);
      if (rc) return rc;
    }
    rc = run_gaussel<T>(pl, A, nx, ny, nx, (long long)nx * ny, pl->th_n, pl->periodic_z, (T)normfft, lam, a, b, c, st);
    if (rc) return rc;
    for (int q = 0; q < nch; ++q) {
      const int k0 = k_lo(q), k1 = k_lo(q + 1), nk = k1 - k0;
      T* pin = p + ((size_t)(k0 + 1) * plane + px + 1);
      T* Aq = A + (size_t)k0 * nx * ny;
      R2RGeom gy{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, nx, nk, ny, 1};
      rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], Aq, Aq, gy, pl->opt.fft_y_lines, st);
      if (rc) return rc;
      R2RGeom gxb{1, 1, nx, px, (long long)nx * ny, px * py, ny, nk, nx, 0};
      rc = run_r2r<T>(ctx, pl->kind[0][1], pl->nt[0], Aq, pin, gxb, pl->opt.fft_x_lines, st);
      if (rc) return rc;
      CK(cudaEventRecord(ctx->chunk_ev[nch + q], st));
      CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->chunk_ev[nch + q], 0));
      CK(cudaMemcpyAsync(hp + (size_t)(k0 + 1) * plane, p + (size_t)(k0 + 1) * plane, (size_t)nk * plane * sizeof(T),
                         cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    CK(cudaStreamSynchronize(ctx->d2h_stream));
    CK(cudaStreamSynchronize(st));
    pl->solves++;
    return 0;
  }
  if (ctx->nranks > 1) {
    const int rcd = ctx->dtdma ? solve_dist_dtdma<T>(pl, p, n, normfft, lam, a, b, c, st)
                               : ((ctx->dist_mode >= 1 && dist_y_fast<T>(pl)) ? solve_dist2<T>(pl, p, n, normfft, lam, a, b, c, pky, ctx->dist_mode == 1, st)
                                                      : solve_dist<T>(pl, p, n, normfft, lam, a, b, c, pky, st));
    if (rcd) return rcd;
    if (mem_kind == CANSB200_MEM_HOST) {
      CK(cudaMemcpyAsync(p_any, ctx->staging.p, nh * sizeof(T), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const int rcs = dist_check(ctx, "solve");
      if (rcs) return rcs;
    }
    pl->solves++;
    return 0;
  }
  T* A = (T*)ctx->scratch.p;
  T* pin = p + (px * py + px + 1);  // p(1,1,1)
  int rc;
  prof_mark(ctx, st);
  // forward x: p (haloed) -> A
  R2RGeom gx{1, 1, px, nx, px * py, (long long)nx * ny, ny, nz, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][0], pl->nt[0], pin, A, gx, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  R2RGeom gy{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, nx, nz, ny, 1};
  // z-major middle stages: B[j][k][i]; the y transforms go A -> B and B -> A, the tridiagonal stage works on B
  bool zm = ctx->zmajor && kind_is_fast(pl->kind[1][0]) && kind_is_fast(pl->kind[1][1]) && pl->nt[1] == ny && !ctx->force_generic;
  if (zm) {
    R2Tables<T>* rt = nullptr;
    int var = ctx->r2_variant[1] < 0 ? 0 : ctx->r2_variant[1];
    if (get_r2_tables<T>(ctx, pl->nt[1], 1, var, &rt) || !rt) zm = false;
  }
  T* B = nullptr;
  if (zm) {
    if (ctx->scratch2.ensure((size_t)nx * ny * nz * sizeof(T))) return fail(CANSB200_ENOMEM, "solve: z-major buffer");
    B = (T*)ctx->scratch2.p;
  }
  const R2RGeom gyf{nx, (long long)nx * nz, 1, 1, (long long)nx * ny, nx, nx, nz, ny, 1};
  const R2RGeom gyb{(long long)nx * nz, nx, 1, 1, nx, (long long)nx * ny, nx, nz, ny, 1};
  // auto: two half-width windows on two streams; the windows' kernels overlap each other's ramp-up / ramp-down
  // (measured -3 % on C3; windows small enough to stay in L2 lose more to launch tails than they gain)
  // ... and with the pivot cache deduplicated in x AND y one full-width tridiagonal launch is better: the four tiles that
  // share pivots run side by side and the sharing happens in L2 (C3: 3.92 ms against 3.97 with the two windows; C4, x only:
  // 9.86 with the windows against 9.95 without)
  const int W = ctx->chain_cols >= 0 ? ctx->chain_cols : ((nx >= 1024 && (nx / 2) % 16 == 0 && !(pl->dx && pl->dy)) ? nx / 2 : 0);
  if (W > 0 && W < nx && thomas_is_pipelined(pl) && !ctx->profiling) {
    // ---- L2-resident chain over x windows, round-robin on auxiliary streams
    ThomasDev<T> D = make_thomas<T>(pl, nx, ny, zm ? (long long)nx * nz : (long long)nx, zm ? (long long)nx : (long long)nx * ny,
                                    pl->th_n, pl->periodic_z, lam, a, b, c);
    rc = gaussel_prepare<T>(pl, D, st);
    if (rc) return rc;
    const int ns = ctx->chain_nstreams;
    while ((int)ctx->aux.size() < ns) {
      cudaStream_t q; cudaEvent_t e;
      CK(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->aux.push_back(q); ctx->aux_done.push_back(e);
    }
    if (!ctx->fork_ev) CK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->fork_ev, st));
    for (int q = 0; q < ns; ++q) CK(cudaStreamWaitEvent(ctx->aux[q], ctx->fork_ev, 0));
    int iw = 0;
    for (int x0 = 0; x0 < nx; x0 += W, ++iw) {
      cudaStream_t sq = ctx->aux[iw % ns];
      const int w = (nx - x0 < W) ? nx - x0 : W;
      R2RGeom gw = zm ? gyf : gy, gwb = zm ? gyb : gy;
      gw.lines_per_group = w;
      gwb.lines_per_group = w;
      rc = run_r2r<T>(ctx, pl->kind[1][0], pl->nt[1], A + x0, (zm ? B : A) + x0, gw, pl->opt.fft_y_lines, sq);
      if (rc) return rc;
      D.xb = x0; D.xn = w;
      rc = gaussel_apply<T>(pl, D, zm ? B : A, (T)normfft, sq);
      if (rc) return rc;
      rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], (zm ? B : A) + x0, A + x0, gwb, pl->opt.fft_y_lines, sq);
      if (rc) return rc;
    }
    for (int q = 0; q < ns; ++q) {
      CK(cudaEventRecord(ctx->aux_done[q], ctx->aux[q]));
      CK(cudaStreamWaitEvent(st, ctx->aux_done[q], 0));
    }
  } else {
    // forward y (in place in A, or A -> z-major B), tridiagonal solve in z, backward y
    rc = run_r2r<T>(ctx, pl->kind[1][0], pl->nt[1], A, zm ? B : A, zm ? gyf : gy, pl->opt.fft_y_lines, st);
    if (rc) return rc;
    prof_mark(ctx, st);
    rc = run_gaussel<T>(pl, zm ? B : A, nx, ny, zm ? (long long)nx * nz : (long long)nx, zm ? (long long)nx : (long long)nx * ny,
                        pl->th_n, pl->periodic_z, (T)normfft, lam, a, b, c, st);
    if (rc) return rc;
    prof_mark(ctx, st);
    rc = run_r2r<T>(ctx, pl->kind[1][1], pl->nt[1], zm ? B : A, A, zm ? gyb : gy, pl->opt.fft_y_lines, st);
    if (rc) return rc;
    prof_mark(ctx, st);
  }
  // backward x
  R2RGeom gxb{1, 1, nx, px, (long long)nx * ny, px * py, ny, nz, nx, 0};
  rc = run_r2r<T>(ctx, pl->kind[0][1], pl->nt[0], A, pin, gxb, pl->opt.fft_x_lines, st);
  if (rc) return rc;
  prof_mark(ctx, st);
  if (mem_kind == CANSB200_MEM_HOST) {
    CK(cudaMemcpyAsync(p_any, ctx->staging.p, nh * sizeof(T), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  pl->solves++;
  return 0;
}

// ---------------------------------------------------------------------------
// gaussel_dtdma (src/solver.f90:309-517) with the z slabs of `nsplit` ranks living on this GPU
template <class T>
static int gaussel_dtdma_impl(cansb200_plan* pl, T* pz, int nx, int ny, int n_rows, int nsplit, const int* starts, int periodic,
                              T norm, const T* lam, const T* a, const T* b, const T* c, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  const size_t ncol = (size_t)nx * ny;
  if (pl->dtdma_big.ensure(3 * ncol * (size_t)n_rows * sizeof(T))) return fail(CANSB200_ENOMEM, "gaussel_dtdma: coefficient arrays");
  if (pl->dtdma_small.ensure((size_t)(11 * nsplit) * ncol * sizeof(T))) return fail(CANSB200_ENOMEM, "gaussel_dtdma: reduced system");
  DtdmaDev<T> D;
  D.nx = nx; D.ny = ny; D.n = n_rows; D.nranks = nsplit; D.periodic = periodic;
  for (int r = 0; r <= nsplit; ++r) D.starts[r] = starts[r];
  D.a = a; D.b = b; D.c = c; D.lam = lam;
  D.st = nullptr; D.slot_big = 0; D.slot_small = 0;   // stage level: no coefficient cache
  T* big = (T*)pl->dtdma_big.p;
  D.Z = big; D.AA = big + ncol * n_rows; D.CC = big + 2 * ncol * n_rows;
  T* sm = (T*)pl->dtdma_small.p;
  D.Z1 = sm;
  D.ra = sm + ncol * nsplit; D.rc = D.ra + 2 * ncol * nsplit; D.rcw = D.rc + 2 * ncol * nsplit;
  D.rp = D.rcw + 2 * ncol * nsplit; D.rp2 = D.rp + 2 * ncol * nsplit;
  const unsigned cb = (unsigned)((ncol + 127) / 128);
  dtdma_coef_kernel<T><<<cb, 128, 0, st>>>(D);
  dtdma_phase1_kernel<T><<<cb, 128, 0, st>>>(D, pz, norm);
  dtdma_reduced_kernel<T><<<cb, 128, 0, st>>>(D);
  dtdma_phase3_kernel<T><<<ctx->num_sms * 8, 256, 0, st>>>(D, pz);
  ctx->launches += 4;
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

int cansb200_gaussel_dtdma(cansb200_plan* pl, void* pz, const int d3[3], int n_rows, int nsplit, const int* starts,
                           int is_periodic, double norm, const void* lam, const void* a, const void* b, const void* c,
                           void* stream) {
  if (!pl || !pz || !d3 || !starts || !a || !b || !c) return fail(CANSB200_EINVAL, "gaussel_dtdma: null argument");
  if (nsplit < 1 || nsplit > CB_DTDMA_MAX_RANKS) return fail(CANSB200_EINVAL, "gaussel_dtdma: nsplit must be 1..16");
  const int nx = d3[0], ny = d3[1], nz = d3[2];
  if (nx < 1 || ny < 1 || n_rows < 1 || n_rows > nz) return fail(CANSB200_EINVAL, "gaussel_dtdma: bad extents");
  if (starts[0] != 0 || starts[nsplit] < n_rows) return fail(CANSB200_EINVAL, "gaussel_dtdma: starts must cover rows 0 .. n_rows-1");
  for (int r = 0; r < nsplit; ++r) {
    const int k1 = starts[r + 1] < n_rows ? starts[r + 1] : n_rows;
    if (k1 - starts[r] < 3) return fail(CANSB200_EINVAL, "gaussel_dtdma: every rank needs at least 3 rows");
  }
  cudaStream_t st = (cudaStream_t)stream;
  return pl->ctx->is_fp32
             ? gaussel_dtdma_impl<float>(pl, (float*)pz, nx, ny, n_rows, nsplit, starts, is_periodic, (float)norm, (const float*)lam,
                                         (const float*)a, (const float*)b, (const float*)c, st)
             : gaussel_dtdma_impl<double>(pl, (double*)pz, nx, ny, n_rows, nsplit, starts, is_periodic, norm, (const double*)lam,
                                          (const double*)a, (const double*)b, (const double*)c, st);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// solver_gaussel_z (src/solver.f90:547-616): tridiagonal solve in z only, no transforms, lambda-less gaussel
template <class T>
static int solve_z_impl(cansb200_plan* pl, void* p_any, const int n[3], double norm, const void* a_any, const void* b_any,
                        const void* c_any, int mem_kind, cudaStream_t st) {
  cansb200_ctx* ctx = pl->ctx;
  const int nx = n[0], ny = n[1], nzl = n[2];
  const long long px = nx + 2, py = ny + 2;
  const size_t nh = (size_t)px * py * (nzl + 2);
  T* p = (T*)p_any;
  const T *a = (const T*)a_any, *b = (const T*)b_any, *c = (const T*)c_any;
  const size_t nlam = (size_t)ctx->n_z[0] * ctx->n_z[1];
  if (!ctx->zero_lam.p) {
    if (ctx->zero_lam.ensure(nlam * sizeof(T))) return fail(CANSB200_ENOMEM, "solve_z: lambda buffer");
    CK(cudaMemsetAsync(ctx->zero_lam.p, 0, nlam * sizeof(T), st));
  }
  if (mem_kind == CANSB200_MEM_HOST) {
    if (ctx->staging.ensure(nh * sizeof(T))) return fail(CANSB200_ENOMEM, "solve_z: staging");
    const size_t nzg = (size_t)ctx->n_z[2];
    if (ctx->coef.ensure((3 * nzg + nlam) * sizeof(T))) return fail(CANSB200_ENOMEM, "solve_z: coefficient staging");
    T* cf = (T*)ctx->coef.p;
    CK(cudaMemcpyAsync(ctx->staging.p, p_any, nh * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf, a_any, nzg * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf + nzg, b_any, nzg * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cf + 2 * nzg, c_any, nzg * sizeof(T), cudaMemcpyHostToDevice, st));
    p = (T*)ctx->staging.p;
    a = cf; b = cf + nzg; c = cf + 2 * nzg;
  }
  const T* lam0 = (const T*)ctx->zero_lam.p;
  T* pin = p + (px * py + px + 1);
  int rc;
  const unsigned long long seq = ++ctx->seq;
  if (ctx->nranks == 1) {
    // z is not decomposed: solve in place on the haloed array (row pitch px, plane pitch px * py)
    if (nzl != ctx->ng[2]) return fail(CANSB200_EINVAL, "solve_z: extents");
    ThomasDev<T> D = make_thomas<T>(pl, nx, ny, px, px * py, pl->th_n, pl->periodic_z, lam0, a, b, c);
    D.nopin = 1;
    rc = gaussel_prepare<T>(pl, D, st);
    if (rc) return rc;
    rc = gaussel_apply<T>(pl, D, pin, (T)norm, st);
    if (rc) return rc;
  } else {
    rc = dist_check(ctx, "solve_z");
    if (rc) return rc;
    const int nyl = ctx->n_z[1], nz = ctx->ng[2];
    T* Cz = (T*)((char*)ctx->region + ctx->off_C);
    const unsigned blocks = (unsigned)(ctx->num_sms * 8);
    ThomasDev<T> D = make_thomas<T>(pl, nx, nyl, nx, (long long)nx * nyl, pl->th_n, pl->periodic_z, lam0, a, b, c);
    D.nopin = 1;
    D.out_rows = (const OutRow<T>*)ctx->ztab.p;
    rc = gaussel_prepare<T>(pl, D, st);   // ahead of the first wait for the peers (see solve_dist)
    if (rc) return rc;
    slab_rows_copy_kernel<T><<<blocks, 256, 0, st>>>(pin, px, px * py, (const DistRow<T>*)ctx->ytab_fwd.p, nx, ny, nzl, 1);
    ctx->launches++;
    CK(cudaGetLastError());
    rc = dist_flag(ctx, CB_SLOT_BAR, seq, true, true, st);
    if (rc) return rc;
    if (thomas_is_pipelined(pl)) {
      rc = gaussel_apply<T>(pl, D, Cz, (T)norm, st);
      if (rc) return rc;
      if (pl->th_n < nz) {
        scatter_rows_kernel<T><<<ctx->num_sms, 256, 0, st>>>(Cz, D.sk, (const DistOutRow<T>*)D.out_rows, pl->th_n, nz, nyl, nx, 0, nx);
        ctx->launches++;
      }
    } else {
      ThomasDev<T> D2 = D;
      D2.out_rows = nullptr;
      rc = gaussel_apply<T>(pl, D2, Cz, (T)norm, st);
      if (rc) return rc;
      scatter_rows_kernel<T><<<ctx->num_sms * 4, 256, 0, st>>>(Cz, D.sk, (const DistOutRow<T>*)D.out_rows, 0, nz, nyl, nx, 0, nx);
      ctx->launches++;
    }
    CK(cudaGetLastError());
    rc = dist_flag(ctx, CB_SLOT_BAR + 1, seq, true, true, st);
    if (rc) return rc;
    slab_rows_copy_kernel<T><<<blocks, 256, 0, st>>>(pin, px, px * py, (const DistRow<T>*)ctx->ytab_bwd.p, nx, ny, nzl, 0);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  if (mem_kind == CANSB200_MEM_HOST) {
    CK(cudaMemcpyAsync(p_any, ctx->staging.p, nh * sizeof(T), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  pl->solves++;
  return 0;
}

extern "C" {

int cansb200_solve_z(cansb200_plan* pl, void* p, const int n[3], int nhalo, double norm, const void* a, const void* b,
                     const void* c, int mem_kind, void* stream) {
  if (!pl || !p || !n || !a || !b || !c) return fail(CANSB200_EINVAL, "solve_z: null argument");
  if (nhalo != 1) return fail(CANSB200_EINVAL, "solve_z: nhalo must be 1");
  cansb200_ctx* ctx = pl->ctx;
  if (ctx->dtdma)
    return fail(CANSB200_EUNSUPPORTED, "solve_z: not available on a CANSB200_CTX_DTDMA context (the reference routes it through gaussel_dtdma, "
                                       "src/solver.f90:592-596; use cansb200_gaussel_dtdma with lambdaxy = NULL on the slab)");
  for (int d = 0; d < 3; ++d)
    if (n[d] != ctx->n[d]) return fail(CANSB200_EINVAL, "solve_z: n differs from the context's local extents");
  if (mem_kind != CANSB200_MEM_HOST && mem_kind != CANSB200_MEM_DEVICE) return fail(CANSB200_EINVAL, "solve_z: bad mem_kind");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ctx_enter(ctx, st);
  if (rc) return rc;
  rc = ctx->is_fp32 ? solve_z_impl<float>(pl, p, n, norm, a, b, c, mem_kind, st)
                    : solve_z_impl<double>(pl, p, n, norm, a, b, c, mem_kind, st);
  const int rl = ctx_leave(ctx, st);
  return rc ? rc : rl;
}

int cansb200_plan_id(cansb200_plan* pl) {
  if (!pl) return 0;
  std::lock_guard<std::mutex> lk(g_plan_mu);
  for (size_t i = 0; i < g_plan_ids.size(); ++i)
    if (g_plan_ids[i] == pl) return (int)i + 1;
  for (size_t i = 0; i < g_plan_ids.size(); ++i)
    if (!g_plan_ids[i]) { g_plan_ids[i] = pl; return (int)i + 1; }
  g_plan_ids.push_back(pl);
  return (int)g_plan_ids.size();
}

cansb200_plan* cansb200_plan_from_id(int id) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  if (id < 1 || id > (int)g_plan_ids.size()) return nullptr;
  return g_plan_ids[id - 1];
}

int cansb200_solve_z_bc(cansb200_ctx* ctx, const char bcz[2], char c_or_f_z, void* p, const int n[3], int nhalo, double norm,
                        const void* a, const void* b, const void* c, int mem_kind, void* stream) {
  if (!ctx || !bcz) return fail(CANSB200_EINVAL, "solve_z_bc: null argument");
  const std::string key{bcz[0], bcz[1], c_or_f_z};
  auto it = ctx->zplans.find(key);
  if (it == ctx->zplans.end()) {
    const char bc6[6] = {'P', 'P', 'P', 'P', bcz[0], bcz[1]};   // x / y kinds are never used by the z-only solve
    const char cf3[3] = {'c', 'c', c_or_f_z};
    cansb200_plan* pl = nullptr;
    const int rc = cansb200_plan_create(ctx, &pl, bc6, cf3, nullptr, nullptr);
    if (rc) return rc;
    it = ctx->zplans.emplace(key, pl).first;
  }
  return cansb200_solve_z(it->second, p, n, nhalo, norm, a, b, c, mem_kind, stream);
}

// ---- the reference's own call sequence: fftini(ng,n_x,n_y,bcxy,c_or_f,arrplan,normfft) ... solver(n,ng,arrplan,normfft,
// lambdaxy,a,b,c,bc,c_or_f,p) ... fftend(arrplan)   (src/fft.f90:25-36,211-217; src/solver_gpu.f90:34-49)
int cansb200_fftini(cansb200_ctx* ctx, const char bcxy[4], const char c_or_f_xy[2], double* normfft_out, int* id_out) {
  if (!ctx || !bcxy || !c_or_f_xy || !id_out) return fail(CANSB200_EINVAL, "fftini: null argument");
  // validate and get normfft from a throw-away plan description (no allocation: only the kind table is evaluated)
  double normfft = 1.0;
  for (int d = 0; d < 2; ++d) {
    const char b0 = bcxy[2 * d], b1 = bcxy[2 * d + 1], cf = c_or_f_xy[d];
    if ((b0 != 'P' && b0 != 'D' && b0 != 'N') || (b1 != 'P' && b1 != 'D' && b1 != 'N') || (cf != 'c' && cf != 'f'))
      return fail(CANSB200_EINVAL, "fftini: bad boundary condition / c_or_f");
    if ((b0 == 'P') != (b1 == 'P')) return fail(CANSB200_EINVAL, "fftini: periodic BCs must come in pairs");
    int kf, kb;
    double n1, n2;
    find_fft(b0, b1, cf, kf, kb, n1, n2);
    const int ii = (b0 == 'D' && b1 == 'D' && cf == 'f') ? 1 : 0;
    normfft *= n1 * (ctx->ng[d] + n2 - ii);
  }
  normfft = 1.0 / normfft;
  if (normfft_out) *normfft_out = ctx->is_fp32 ? (double)(1.0f / (float)(1.0 / normfft)) : normfft;
  cansb200_ctx::FftiniRec rec;
  memcpy(rec.bcxy, bcxy, 4);
  memcpy(rec.cf, c_or_f_xy, 2);
  rec.live = true;
  for (size_t i = 0; i < ctx->fftini_recs.size(); ++i)
    if (!ctx->fftini_recs[i].live) { ctx->fftini_recs[i] = rec; *id_out = (int)i + 1; return 0; }
  ctx->fftini_recs.push_back(rec);
  *id_out = (int)ctx->fftini_recs.size();
  return 0;
}

int cansb200_fftend(cansb200_ctx* ctx, int id) {
  if (!ctx || id < 1 || id > (int)ctx->fftini_recs.size() || !ctx->fftini_recs[id - 1].live) return fail(CANSB200_EINVAL, "fftend: unknown id");
  auto& rec = ctx->fftini_recs[id - 1];
  for (auto& kv : rec.byz) cansb200_plan_destroy(kv.second);
  rec.byz.clear();
  rec.live = false;
  return 0;
}

// the plan behind (fftini id, z boundary conditions, c_or_f(3), eigenvalue order), created on first use
static int solver_plan(cansb200_ctx* ctx, int id, const char bc[6], const char c_or_f[3], int lambda_order, cansb200_plan** out) {
  if (!ctx || !bc || !c_or_f) return fail(CANSB200_EINVAL, "solver: null argument");
  if (id < 1 || id > (int)ctx->fftini_recs.size() || !ctx->fftini_recs[id - 1].live) return fail(CANSB200_EINVAL, "solver: unknown fftini id");
  auto& rec = ctx->fftini_recs[id - 1];
  if (memcmp(rec.bcxy, bc, 4) != 0 || rec.cf[0] != c_or_f[0] || rec.cf[1] != c_or_f[1])
    return fail(CANSB200_EINVAL, "solver: x / y boundary conditions differ from the ones fftini was called with");
  const std::string key{bc[4], bc[5], c_or_f[2], (char)('0' + (lambda_order ? 1 : 0))};
  auto it = rec.byz.find(key);
  if (it == rec.byz.end()) {
    cansb200_options o;
    for (int* q = (int*)&o; q < (int*)(&o + 1); ++q) *q = -1;
    o.cache_slots = 3;   // one per RK sub-step: the Helmholtz solves of a variable see three different alpha
    o.lambda_order = lambda_order ? 1 : 0;
    cansb200_plan* pl = nullptr;
    const int rc = cansb200_plan_create(ctx, &pl, bc, c_or_f, &o, nullptr);
    if (rc) return rc;
    it = rec.byz.emplace(key, pl).first;
  }
  *out = it->second;
  return 0;
}

int cansb200_solver(cansb200_ctx* ctx, int id, const char bc[6], const char c_or_f[3], void* p, const int n[3], int nhalo,
                    double normfft, const void* lambdaxy, const void* a, const void* b, const void* c, int lambda_order,
                    int mem_kind, void* stream) {
  cansb200_plan* pl = nullptr;
  const int rc = solver_plan(ctx, id, bc, c_or_f, lambda_order, &pl);
  if (rc) return rc;
  return cansb200_solve(pl, p, n, nhalo, normfft, lambdaxy, a, b, c, mem_kind, stream);
}

int cansb200_solver_fillps(cansb200_ctx* ctx, int id, const char bc[6], const char c_or_f[3], void* p, const int n[3], int nhalo,
                           double normfft, const void* lambdaxy, const void* a, const void* b, const void* c, int lambda_order,
                           const double dli[3], const void* dzfi, double dti, const void* u, const void* v, const void* w,
                           const int is_bound[6], const int have[3], const double rhsb[6], void* stream) {
  cansb200_plan* pl = nullptr;
  const int rc = solver_plan(ctx, id, bc, c_or_f, lambda_order, &pl);
  if (rc) return rc;
  return cansb200_solve_fillps(pl, p, n, nhalo, normfft, lambdaxy, a, b, c, dli, dzfi, dti, u, v, w, is_bound, have, rhsb, stream);
}

// planes and values of updt_rhs_b's six wall terms (src/bound.f90:514-598)
static int rhsb_planes(const cansb200_ctx* ctx, const char cf[3], const char bc[6], const int n[3], const int is_bound[6],
                       const int have[3], const double rhsb[6], double norm, RhsbPlanes& B) {
  for (int d = 0; d < 3; ++d) {
    const int q = (cf[d] == 'f' && bc[2 * d + 1] == 'D') ? 1 : 0;   // src/bound.f90:528-530
    for (int sd = 0; sd < 2; ++sd) {
      const bool on = have[d] && is_bound[2 * d + sd];
      B.idx[d][sd] = on ? (sd == 0 ? 1 : n[d] - q) : 0;
      // value * norm in the working precision, as `rhsbx(j,k,0)*norm` is evaluated
      B.val[d][sd] = ctx->is_fp32 ? (double)((float)rhsb[2 * d + sd] * (float)norm) : rhsb[2 * d + sd] * norm;
      if (on && B.idx[d][sd] < 1) return fail(CANSB200_EINVAL, "updt_rhs_b: empty direction");
    }
  }
  return 0;
}

int cansb200_updt_rhs_b(cansb200_ctx* ctx, const char cf[3], const char bc[6], const int n[3], const int is_bound[6],
                        const int have[3], const double rhsb[6], double norm, void* p, void* stream) {
  if (!ctx || !cf || !bc || !n || !is_bound || !have || !rhsb || !p) return fail(CANSB200_EINVAL, "updt_rhs_b: null argument");
  RhsbPlanes B;
  const int rcb = rhsb_planes(ctx, cf, bc, n, is_bound, have, rhsb, norm, B);
  if (rcb) return rcb;
  cudaStream_t st = (cudaStream_t)stream;
  for (int d = 0; d < 3; ++d) {
    if (!B.idx[d][0] && !B.idx[d][1]) continue;
    const long long face = d == 0 ? (long long)n[1] * n[2] : (d == 1 ? (long long)n[0] * n[2] : (long long)n[0] * n[1]);
    const unsigned blocks = (unsigned)((face + 255) / 256 < 1184 ? (face + 255) / 256 : 1184);
    if (ctx->is_fp32) updt_rhs_b_kernel<float><<<blocks, 256, 0, st>>>((float*)p, n[0], n[1], n[2], B, d);
    else updt_rhs_b_kernel<double><<<blocks, 256, 0, st>>>((double*)p, n[0], n[1], n[2], B, d);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int cansb200_solve(cansb200_plan* pl, void* p, const int n[3], int nhalo, double normfft, const void* lambdaxy,
                   const void* a, const void* b, const void* c, int mem_kind, void* stream) {
  if (!pl || !p || !n || !lambdaxy || !a || !b || !c) return fail(CANSB200_EINVAL, "solve: null argument");
  if (nhalo != 1) return fail(CANSB200_EINVAL, "solve: nhalo must be 1");
  cansb200_ctx* ctx = pl->ctx;
  for (int d = 0; d < 3; ++d)
    if (n[d] != ctx->n[d]) return fail(CANSB200_EINVAL, "solve: n differs from the context's local extents");
  if (mem_kind != CANSB200_MEM_HOST && mem_kind != CANSB200_MEM_DEVICE) return fail(CANSB200_EINVAL, "solve: bad mem_kind");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ctx_enter(ctx, st);
  if (rc) return rc;
  rc = ctx->is_fp32 ? solve_impl<float>(pl, p, n, normfft, lambdaxy, a, b, c, mem_kind, st)
                    : solve_impl<double>(pl, p, n, normfft, lambdaxy, a, b, c, mem_kind, st);
  const int rl = ctx_leave(ctx, st);
  return rc ? rc : rl;
}

int cansb200_solve_fillps(cansb200_plan* pl, void* p, const int n[3], int nhalo, double normfft, const void* lambdaxy,
                          const void* a, const void* b, const void* c, const double dli[3], const void* dzfi, double dti,
                          const void* u, const void* v, const void* w, const int is_bound[6], const int have[3],
                          const double rhsb[6], void* stream) {
  if (!pl || !p || !n || !lambdaxy || !a || !b || !c || !dli || !dzfi || !u || !v || !w)
    return fail(CANSB200_EINVAL, "solve_fillps: null argument");
  if (nhalo != 1) return fail(CANSB200_EINVAL, "solve_fillps: nhalo must be 1");
  cansb200_ctx* ctx = pl->ctx;
  for (int d = 0; d < 3; ++d)
    if (n[d] != ctx->n[d]) return fail(CANSB200_EINVAL, "solve_fillps: n differs from the context's local extents");
  const bool walls = is_bound && have && rhsb;
  if (!walls && (is_bound || have || rhsb)) return fail(CANSB200_EINVAL, "solve_fillps: is_bound, have and rhsb come together or not at all");
  cansb200_ctx::FuseSrc fs;
  fs.u = u; fs.v = v; fs.w = w; fs.dzfi = dzfi;
  fs.dti = dti; fs.dxi = dli[0]; fs.dyi = dli[1];
  fs.any_rhsb = 0;
  for (int d = 0; d < 3; ++d)
    for (int sd = 0; sd < 2; ++sd) { fs.B.idx[d][sd] = 0; fs.B.val[d][sd] = 0.0; }
  if (walls) {
    const int rcb = rhsb_planes(ctx, pl->cf, pl->bc, n, is_bound, have, rhsb, 1.0, fs.B);
    if (rcb) return rcb;
    for (int d = 0; d < 3; ++d) fs.any_rhsb |= (fs.B.idx[d][0] | fs.B.idx[d][1]) ? 1 : 0;
  }
  const bool fuse = ctx->is_fp32 ? can_fuse_fillps<float>(pl) : can_fuse_fillps<double>(pl);
  pl->last_fused = fuse ? 1 : 0;
  if (!fuse) {
    // lengths / kinds the two-for-one x kernels do not serve: the three steps one after the other (all of them CUDA)
    int rc = cansb200_fillps(ctx, n, dli, dzfi, dti, u, v, w, p, stream);
    if (rc) return rc;
    if (walls && fs.any_rhsb) {
      rc = cansb200_updt_rhs_b(ctx, pl->cf, pl->bc, n, is_bound, have, rhsb, 1.0, p, stream);
      if (rc) return rc;
    }
    return cansb200_solve(pl, p, n, nhalo, normfft, lambdaxy, a, b, c, CANSB200_MEM_DEVICE, stream);
  }
  const size_t esz = ctx->esz;
  fs.pin = (const char*)p + ((size_t)(n[0] + 2) * (n[1] + 2) + (size_t)(n[0] + 2) + 1) * esz;   // p(1,1,1)
  // u, v, w as the kernel addresses them: element (1,1,1) of arrays shaped like p
  const size_t o111 = ((size_t)(n[0] + 2) * (n[1] + 2) + (size_t)(n[0] + 2) + 1) * esz;
  fs.u = (const char*)u + o111; fs.v = (const char*)v + o111; fs.w = (const char*)w + o111;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ctx_enter(ctx, st);
  if (rc) return rc;
  ctx->fuse = &fs;
  rc = ctx->is_fp32 ? solve_impl<float>(pl, p, n, normfft, lambdaxy, a, b, c, CANSB200_MEM_DEVICE, st)
                    : solve_impl<double>(pl, p, n, normfft, lambdaxy, a, b, c, CANSB200_MEM_DEVICE, st);
  ctx->fuse = nullptr;
  const int rl = ctx_leave(ctx, st);
  return rc ? rc : rl;
}

int cansb200_r2r(cansb200_ctx* ctx, int kind, int nt, int axis, void* arr, const int d3[3], void* stream) {
  if (!ctx || !arr || !d3) return fail(CANSB200_EINVAL, "r2r: null argument");
  if (axis != 0 && axis != 1) return fail(CANSB200_EINVAL, "r2r: axis must be 0 (x) or 1 (y)");
  const int nx = d3[0], ny = d3[1], nz = d3[2];
  if (nt < 1 || nt > (axis == 0 ? nx : ny)) return fail(CANSB200_EINVAL, "r2r: bad transform length");
  const bool known = kind == K_R2HC || kind == K_HC2R || (kind >= K_REDFT00 && kind <= K_RODFT11);
  if (!known) return fail(CANSB200_EINVAL, "r2r: unknown transform kind");
  R2RGeom g = axis == 0 ? R2RGeom{1, 1, nx, nx, (long long)nx * ny, (long long)nx * ny, ny, nz, nx, 0}
                        : R2RGeom{nx, nx, 1, 1, (long long)nx * ny, (long long)nx * ny, nx, nz, ny, 1};
  cudaStream_t st = (cudaStream_t)stream;
  return ctx->is_fp32 ? run_r2r<float>(ctx, kind, nt, (const float*)arr, (float*)arr, g, 0, st)
                      : run_r2r<double>(ctx, kind, nt, (const double*)arr, (double*)arr, g, 0, st);
}

int cansb200_gaussel(cansb200_plan* pl, void* pz, const int d3[3], int n_rows, int is_periodic, double norm,
                     const void* lam, const void* a, const void* b, const void* c, void* stream) {
  if (!pl || !pz || !d3 || !lam || !a || !b || !c) return fail(CANSB200_EINVAL, "gaussel: null argument");
  const int nx = d3[0], ny = d3[1];
  if (pl->ctx->dtdma) return fail(CANSB200_EUNSUPPORTED, "gaussel: a CANSB200_CTX_DTDMA context keeps no pivot cache (use cansb200_gaussel_dtdma)");
  if (nx != pl->ctx->n_z[0] || ny != pl->ctx->n_z[1]) return fail(CANSB200_EINVAL, "gaussel: extents differ from the plan's");
  cudaStream_t st = (cudaStream_t)stream;
  return pl->ctx->is_fp32
             ? run_gaussel<float>(pl, (float*)pz, nx, ny, nx, (long long)nx * ny, n_rows, is_periodic, (float)norm,
                                  (const float*)lam, (const float*)a, (const float*)b, (const float*)c, st)
             : run_gaussel<double>(pl, (double*)pz, nx, ny, nx, (long long)nx * ny, n_rows, is_periodic, norm,
                                   (const double*)lam, (const double*)a, (const double*)b, (const double*)c, st);
}

int cansb200_fill_hash(cansb200_ctx* ctx, void* p, const int n[3], const int lo[3], int nhalo, unsigned long long seed,
                       void* stream) {
  if (!ctx || !p || !n || !lo) return fail(CANSB200_EINVAL, "fill_hash: null argument");
  if (nhalo != 0 && nhalo != 1) return fail(CANSB200_EINVAL, "fill_hash: nhalo must be 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  const long long tot = (long long)(n[0] + 2 * nhalo) * (n[1] + 2 * nhalo) * (n[2] + 2 * nhalo);
  const unsigned blocks = (unsigned)((tot + 255) / 256 < 65535LL * 16 ? (tot + 255) / 256 : 65535LL * 16);
  if (ctx->is_fp32)
    fill_hash_kernel<float><<<blocks, 256, 0, st>>>((float*)p, n[0], n[1], n[2], lo[0] - 1, lo[1] - 1, lo[2] - 1, ctx->ng[0], ctx->ng[1], nhalo, seed);
  else
    fill_hash_kernel<double><<<blocks, 256, 0, st>>>((double*)p, n[0], n[1], n[2], lo[0] - 1, lo[1] - 1, lo[2] - 1, ctx->ng[0], ctx->ng[1], nhalo, seed);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

int cansb200_fillps(cansb200_ctx* ctx, const int n[3], const double dli[3], const void* dzfi, double dti, const void* u,
                    const void* v, const void* w, void* p, void* stream) {
  if (!ctx || !n || !dli || !dzfi || !u || !v || !w || !p) return fail(CANSB200_EINVAL, "fillps: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g3, b3;
  if (ctx->aux_3d && aux_geom(n[0], n[1], n[2], g3, b3)) {
    if (ctx->is_fp32)
      fillps3d_kernel<float><<<g3, b3, 0, st>>>(n[0], n[1], (float)dli[0], (float)dli[1], (const float*)dzfi, (float)dti, (const float*)u,
                                                (const float*)v, (const float*)w, (float*)p);
    else
      fillps3d_kernel<double><<<g3, b3, 0, st>>>(n[0], n[1], dli[0], dli[1], (const double*)dzfi, dti, (const double*)u, (const double*)v,
                                                 (const double*)w, (double*)p);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  const long long tot = (long long)n[0] * n[1] * n[2];
  const unsigned blocks = (unsigned)((tot + 255) / 256);
  if (ctx->is_fp32)
    fillps_kernel<float><<<blocks, 256, 0, st>>>(n[0], n[1], n[2], (float)dli[0], (float)dli[1], (const float*)dzfi, (float)dti,
                                                 (const float*)u, (const float*)v, (const float*)w, (float*)p);
  else
    fillps_kernel<double><<<blocks, 256, 0, st>>>(n[0], n[1], n[2], dli[0], dli[1], (const double*)dzfi, dti, (const double*)u,
                                                  (const double*)v, (const double*)w, (double*)p);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

int cansb200_correc(cansb200_ctx* ctx, const int n[3], const double dli[3], const void* dzci, double dt, const void* p, void* u,
                    void* v, void* w, void* stream) {
  if (!ctx || !n || !dli || !dzci || !u || !v || !w || !p) return fail(CANSB200_EINVAL, "correc: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g3, b3;
  if (ctx->aux_3d && aux_geom(n[0] + 2, n[1] + 2, n[2] + 2, g3, b3)) {
    if (ctx->is_fp32)
      correc3d_kernel<float><<<g3, b3, 0, st>>>(n[0], n[1], n[2], (float)dli[0], (float)dli[1], (const float*)dzci, (float)dt, (const float*)p,
                                                (float*)u, (float*)v, (float*)w);
    else
      correc3d_kernel<double><<<g3, b3, 0, st>>>(n[0], n[1], n[2], dli[0], dli[1], (const double*)dzci, dt, (const double*)p, (double*)u,
                                                 (double*)v, (double*)w);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  const long long tot = (long long)(n[0] + 2) * (n[1] + 2) * (n[2] + 2);
  const unsigned blocks = (unsigned)((tot + 255) / 256);
  if (ctx->is_fp32)
    correc_kernel<float><<<blocks, 256, 0, st>>>(n[0], n[1], n[2], (float)dli[0], (float)dli[1], (const float*)dzci, (float)dt,
                                                 (const float*)p, (float*)u, (float*)v, (float*)w);
  else
    correc_kernel<double><<<blocks, 256, 0, st>>>(n[0], n[1], n[2], dli[0], dli[1], (const double*)dzci, dt, (const double*)p,
                                                  (double*)u, (double*)v, (double*)w);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

int cansb200_chkdiv(cansb200_ctx* ctx, const int n[3], const double dli[3], const void* dzfi, const void* u, const void* v,
                    const void* w, double* divtot_sum, double* divmax, void* stream) {
  if (!ctx || !n || !dli || !dzfi || !u || !v || !w || !divtot_sum || !divmax) return fail(CANSB200_EINVAL, "chkdiv: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* res = nullptr;
  CK(cudaMalloc(&res, 2 * sizeof(double)));
  CK(cudaMemsetAsync(res, 0, 2 * sizeof(double), st));
  const long long tot = (long long)n[0] * n[1] * n[2];
  unsigned blocks = (unsigned)((tot + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (ctx->is_fp32)
    chkdiv_kernel<float><<<blocks, 256, 0, st>>>(n[0], n[1], n[2], (float)dli[0], (float)dli[1], (const float*)dzfi, (const float*)u,
                                                 (const float*)v, (const float*)w, res);
  else
    chkdiv_kernel<double><<<blocks, 256, 0, st>>>(n[0], n[1], n[2], dli[0], dli[1], (const double*)dzfi, (const double*)u,
                                                  (const double*)v, (const double*)w, res);
  ctx->launches++;
  double h[2];
  cudaError_t e = cudaMemcpyAsync(h, res, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(res);
  if (e != cudaSuccess) return fail(CANSB200_ECUDA, std::string("chkdiv: ") + cudaGetErrorString(e));
  *divtot_sum = h[0];
  *divmax = h[1];
  return 0;
}

}  // extern "C"
