/* cans_b200 -- C ABI of the B200-native FFT-based Poisson/Helmholtz solver.
 *
 * Drop-in boundary for the ONE hot path of CaNS: `solver` / `solver_gpu`
 * (reference /root/reference/src/solver.f90:17-112, src/solver_gpu.f90:34-276),
 * reused by `solve_helmholtz` (src/solve_helmholtz.f90:28-75).  The Fortran
 * modules of `fortran/solver_b200.f90` (mod_fft, mod_solver / mod_solver_gpu,
 * mod_workspaces) bind these symbols through ISO_C_BINDING so that
 * main.f90 / rk.f90 / mom.f90 / initsolver.f90 / solve_helmholtz.f90 stay unchanged.
 *
 * Conventions
 *  - every function returns 0 on success, a negative CANSB200_E* code otherwise,
 *    and never aborts (the reference `error stop`s: src/fft.f90:525,696 -- the
 *    Fortran shim turns a non-zero status into `error stop`);
 *  - fields are Fortran arrays p(0:n1+1,0:n2+1,0:n3+1), x fastest, one ghost
 *    cell per side (src/main.f90:164-168); only the interior is written;
 *  - `mem_kind` says where p / lambdaxy / a / b / c live (host or device);
 *  - all work is stream-ordered on the caller's CUDA stream (the reference
 *    enqueues everything on OpenACC queue 1: src/workspaces.f90:101-106);
 *    no host synchronisation happens inside cansb200_solve in device mode;
 *  - one solve at a time per context (its plans share scratch, staging and the exchange region): one host thread issues
 *    work, as in the reference; solves enqueued on different streams are ordered against each other by the library;
 *  - precision is fixed per context: FP64, or FP32 when CaNS is built with
 *    -D_SINGLE_PRECISION (src/types.f90:12-18).
 */
#ifndef CANS_B200_H
#define CANS_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cansb200_ctx cansb200_ctx;
typedef struct cansb200_plan cansb200_plan;

enum {
  CANSB200_OK = 0,
  CANSB200_EINVAL = -1,   /* bad argument */
  CANSB200_ECUDA = -2,    /* CUDA runtime error, see cansb200_last_error() */
  CANSB200_ENOMEM = -3,
  CANSB200_EUNSUPPORTED = -4,
  CANSB200_ECOMM = -5     /* NCCL / peer-memory error */
};
enum { CANSB200_MEM_HOST = 0, CANSB200_MEM_DEVICE = 1 };

/* Tunables; negative = library default.  Read once by cansb200_plan_create. */
typedef struct cansb200_options {
  int thomas_variant;   /* 0 = sequential two-sweep (any nz); 1 = pipelined chunk-parallel on-chip (nz <= 1024; above 512 rows a cluster
                           of two CTAs shares each 16-column tile); 2 = as 1 but 8-column tiles instead of clusters above 512 rows;
                           3 = as 1 but the tiles are fetched with cp.async instead of TMA (cp.async.bulk.tensor) */
  int cache_slots;      /* factorisation cache entries per plan (1..8); Helmholtz plans want 3 */
  int fft_x_lines;      /* lines per tile of the contiguous transforms (0 = auto) */
  int fft_y_lines;      /* 8 or 16: x-width of the strided-transform tile (0 = auto) */
  int exchange;         /* reserved (multi-GPU exchange flavour): only the fused peer stores over CUDA IPC exist; the value is ignored */
  int lambda_order;     /* order of lambdaxy along a periodic direction: 0 (default) = FFTW halfcomplex (r0, r1, .., r[n/2], i[n/2-1], .., i1),
                           what the CPU build's initsolver produces; 1 = the packed order (r0, r[n/2], r1, i1, r2, i2, ..)
                           an _OPENACC build's initsolver produces (src/initsolver.f90:98-117), so that initsolver stays
                           unchanged on an OpenACC host too.  On several ranks the rows of a periodic y direction are then
                           dealt out to the z pencils in packed order, i.e. lambdaxy(lo_z(1):hi_z(1), lo_z(2):hi_z(2)) means what
                           it means in the reference */
  int pivot_dedup;      /* pivot cache of the tridiagonal stage: -1 / 1 (default) = keep one copy of the pivots of the columns that share
                           their eigenvalue (the real and the imaginary part of a mode in a periodic direction: lambda(i) = lambda(n - i)),
                           i.e. half of the cache and of its HBM stream per periodic direction (FP64, pipelined variants).  initsolver's
                           two halves agree to rounding (<= 2e-11 relative), which moves the pressure by <= 1e-14 in relative L2; the
                           symmetry of the caller's lambdaxy is verified on the device to 1e-10 (the first solve falls back to the
                           full cache if it does not hold);  0 = always the full cache */
  int tall_tiles;       /* one-GPU solves of shallow grids (nz <= 256): -1 / 1 (default) = the tridiagonal stage solves several y rows per
                           tile (the z-major field is contiguous in (j, k)), which keeps ~512 rows per tile; 0 = one y row per tile */
  int reserved[8];
} cansb200_options;

/* -- context: replaces initmpi's cuDecomp setup (src/initmpi.f90:84-146), common_cudecomp.f90
 *    and workspaces.f90's init_wspace_arrays.  ng = global grid, dims = processor grid
 *    (only dims = [1, nranks] is implemented: z slabs, which are both the x pencil and the y pencil of that processor
 *    grid, so ipencil_axis may be 1 or 2; on one rank it may also be 3).  nccl_id is accepted for signature
 *    parity with cuDecomp's rendezvous (cudecomp.cc:65-69) and ignored: the exchange uses CUDA IPC peer
 *    mappings, see cansb200_dist_export / cansb200_dist_connect below. */
int cansb200_init(cansb200_ctx** ctx, const int ng[3], const int dims[2], int ipencil_axis,
                  int rank, int nranks, const void* nccl_id, int is_fp32);
int cansb200_finalize(cansb200_ctx* ctx);

/* -- multi-GPU rendezvous (one rank per GPU, all on one NVLink box).  The transposes of the reference
 *    (cudecompTranspose{YtoZ,ZtoY}: src/solver_gpu.f90:157,167) are peer-mapped stores issued by the
 *    producing kernels, so each rank must map its peers' exchange regions once:
 *      1. every rank calls cansb200_dist_export into a blob of cansb200_dist_blob_size() bytes,
 *      2. the host gathers the blobs of all ranks in rank order (MPI_Allgather / torch.distributed),
 *      3. every rank calls cansb200_dist_connect with the gathered array.
 *    A device-side wait for a peer that does not show up within 20 s gives up instead of hanging the GPU and sets a
 *    sticky status word: every later cansb200_solve / _solve_z on the context (and the host-memory solve that saw it)
 *    returns CANSB200_ECOMM; cansb200_dist_status reads the word. */
int cansb200_dist_blob_size(void);
int cansb200_dist_export(cansb200_ctx* ctx, void* blob);
int cansb200_dist_connect(cansb200_ctx* ctx, const void* blobs);
int cansb200_dist_status(cansb200_ctx* ctx, int* status);
/* The same rendezvous when all ranks are contexts of ONE process on ONE device (ctxs[r] = the context of rank r):
 * the peers' regions are plain device pointers.  Every rank's solve must then be enqueued on its own stream before
 * the host synchronises (the device-side waits of one rank need the kernels of the others to run; for the same reason
 * the process must load its kernels eagerly, CUDA_MODULE_LOADING=EAGER, or have run every kernel once).  It exists so that
 * a single-GPU box exercises the kernels, row tables and flags of the multi-GPU path (tests/test_gpu_parity.py). */
int cansb200_dist_connect_local(cansb200_ctx* const* ctxs, int n);

/* local extents of this rank: n = x-pencil (nx, ny, nz/P) as main.f90 sees it, lo_z / n_z =
 * z-pencil (nx, ny/P, nz) that lambdaxy is indexed by (src/initsolver.f90:54-58). 1-based lo. */
int cansb200_get_extents(const cansb200_ctx* ctx, int n[3], int lo[3], int n_z[3], int lo_z[3]);

/* -- plan: replaces fftini / fftend (src/fft.f90:25-245).  bc = {x0,x1,y0,y1,z0,z1} in
 *    {'P','D','N'}, c_or_f = {'c'|'f'} per direction.  normfft_out receives fftini's normfft. */
int cansb200_plan_create(cansb200_ctx* ctx, cansb200_plan** plan, const char bc[6],
                         const char c_or_f[3], const cansb200_options* opt, double* normfft_out);
int cansb200_plan_destroy(cansb200_plan* plan);

/* -- the solve: replaces solver / solver_gpu.  Same argument meaning as the reference:
 *    n = local interior extents, nhalo = 1, normfft (normfft/alpha for Helmholtz),
 *    lambdaxy(n_z(1),n_z(2)), a/b/c(n_z(3)) -- b already shifted by 1/alpha by the caller
 *    (src/solve_helmholtz.f90:63-71).  p is solved in place. */
int cansb200_solve(cansb200_plan* plan, void* p, const int n[3], int nhalo, double normfft,
                   const void* lambdaxy, const void* a, const void* b, const void* c,
                   int mem_kind, void* stream);

/* -- z-only solve: replaces solver_gaussel_z / solver_gaussel_z_gpu (src/solver.f90:547-616,
 *    src/solver_gpu.f90:956-1105; implicit z diffusion, is_impdiff_1d): the lambda-less gaussel on the interior of
 *    the haloed p, no transforms.  The plan supplies the z variant (periodic closure, q = 1 for a face-centred
 *    Dirichlet top).  On a z-decomposed grid the slab travels to the z pencils and back over the same
 *    peer-mapped row tables as the full solve.  norm as in the reference (1 for the diffusion solves). */
int cansb200_solve_z(cansb200_plan* plan, void* p, const int n[3], int nhalo, double norm,
                     const void* a, const void* b, const void* c, int mem_kind, void* stream);

/* The same solve for a host that holds no plan for it: the reference's call is
 *   solver_gaussel_z(n,ng,hi,a,b,c,bcz,c_or_f,norm,p)     (src/solve_helmholtz.f90:73, src/solver_gpu.f90:956-963)
 * -- no arrplan.  The context keeps one z-only plan per (bcz, c_or_f(3)), created on first use. */
int cansb200_solve_z_bc(cansb200_ctx* ctx, const char bcz[2], char c_or_f_z, void* p, const int n[3], int nhalo, double norm,
                        const void* a, const void* b, const void* c, int mem_kind, void* stream);

/* -- the reference's own call sequence, argument for argument (what fortran/solver_b200.f90 binds):
 *      call fftini(ng,n_x,n_y,bcxy,c_or_f,arrplan,normfft)                      src/fft.f90:25-36
 *      call solver(n,ng,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p)              src/solver_gpu.f90:34-49
 *      call fftend(arrplan)                                                      src/fft.f90:211-217
 *    fftini sees the x / y boundary conditions only (bcxy = {x0,x1,y0,y1}), so cansb200_fftini just records them and
 *    returns normfft and a small positive integer id -- which fits arrplan(1,1) on every build of the reference, the CUDA
 *    one included (`integer :: arrplan(2,2)`).  cansb200_solver creates the actual plan on first use, one per (id, z
 *    boundary conditions, c_or_f(3)) with three pivot-cache slots, and solves.  lambda_order = 1 on an _OPENACC host
 *    (cansb200_options::lambda_order). */
int cansb200_fftini(cansb200_ctx* ctx, const char bcxy[4], const char c_or_f_xy[2], double* normfft_out, int* id_out);
int cansb200_fftend(cansb200_ctx* ctx, int id);
int cansb200_solver(cansb200_ctx* ctx, int id, const char bc[6], const char c_or_f[3], void* p, const int n[3], int nhalo,
                    double normfft, const void* lambdaxy, const void* a, const void* b, const void* c, int lambda_order,
                    int mem_kind, void* stream);

/* -- integer names for plans.  On the reference's CUDA build `arrplan` is `integer, dimension(2,2)` (cuFFT handles:
 *    src/main.f90:94,107, src/fft.f90:31-35), not type(C_PTR): a 64-bit pointer does not fit one element.  The shim stores
 *    cansb200_plan_id(plan) (a small positive integer, unique per process while the plan lives) in arrplan(1,1) and gets
 *    the plan back with cansb200_plan_from_id (NULL for an unknown id). */
int cansb200_plan_id(cansb200_plan* plan);
cansb200_plan* cansb200_plan_from_id(int id);

/* -- updt_rhs_b (src/bound.f90:514-598) on the device: adds rhsb*(d, side) * norm to the first / last interior planes.
 *    rhsb[d][side] are the (uniform) wall values of bc_rhs (src/initsolver.f90:189-232); is_bound[d][side] != 0 where this
 *    rank owns the wall; have[d] = 0 skips a direction (an absent optional argument); norm = alpha for Helmholtz solves,
 *    1 otherwise.  c_or_f / bc as in cansb200_plan_create (face-centred Dirichlet tops end one plane early). */
int cansb200_updt_rhs_b(cansb200_ctx* ctx, const char c_or_f[3], const char bc[6], const int n[3], const int is_bound[6],
                        const int have[3], const double rhsb[6], double norm, void* p, void* stream);

/* -- stage-level entry points (device pointers only), so that tests can compare every stage
 *    with the oracle the way the reference composes them:
 *    cansb200_r2r     == one `call fft(arrplan(idir_fb), arr)`            (src/fft.f90:247-258)
 *    cansb200_gaussel == `call gaussel(nx,ny,n,0,a,b,c,is_periodic,norm,p,lambdaxy)` (src/solver.f90:114) */
int cansb200_r2r(cansb200_ctx* ctx, int kind, int n_transform, int axis /*0=x,1=y*/,
                 void* arr, const int dims3[3] /*nx,ny,nz of the haloless array*/, void* stream);
int cansb200_gaussel(cansb200_plan* plan, void* pz, const int dims3[3], int n_rows, int is_periodic,
                     double norm, const void* lambdaxy, const void* a, const void* b, const void* c,
                     void* stream);

/* -- distributed TDMA, stage level (device pointers): the arithmetic of gaussel_dtdma / gaussel_dtdma_gpu
 *    (src/solver.f90:309-517, src/solver_gpu.f90:430-695; is_poisson_dtdma) on a haloless field pz[k][j][i] whose
 *    rows are split at `starts` (nsplit + 1 entries) into the z slabs of nsplit ranks, all living on this GPU:
 *    inner-row elimination per slab, reduced 2-rows-per-rank system, update.  lambdaxy may be NULL (z-only
 *    variant).  This entry point keeps all slabs on one GPU (stage-level parity with the oracle); on a CANSB200_CTX_DTDMA
 *    context cansb200_solve runs the same elimination per rank and gathers the reduced rows over the peer mappings. */
int cansb200_gaussel_dtdma(cansb200_plan* plan, void* pz, const int dims3[3], int n_rows, int nsplit, const int* starts,
                           int is_periodic, double norm, const void* lambdaxy, const void* a, const void* b,
                           const void* c, void* stream);

/* -- the steps either side of the path (device pointers, haloed arrays):
 *    fillps (src/fillps.f90:13-51), correc (src/correc.f90:13-60), chkdiv (src/chkdiv.f90:15-54) */
int cansb200_fillps(cansb200_ctx* ctx, const int n[3], const double dli[3], const void* dzfi, double dti,
                    const void* u, const void* v, const void* w, void* p, void* stream);
int cansb200_correc(cansb200_ctx* ctx, const int n[3], const double dli[3], const void* dzci, double dt,
                    const void* p, void* u, void* v, void* w, void* stream);
int cansb200_chkdiv(cansb200_ctx* ctx, const int n[3], const double dli[3], const void* dzfi,
                    const void* u, const void* v, const void* w, double* divtot_sum, double* divmax,
                    void* stream);

/* -- the pressure-correction right-hand side fused into the solve (device pointers): one call for
 *      call fillps(n,dli,dzfi,dtrki,u,v,w,pp)                                          src/main.f90:465, src/fillps.f90:13-51
 *      call updt_rhs_b(['c','c','c'],cbcpre,n,is_bound,rhsbp%x,rhsbp%y,rhsbp%z,pp)     src/main.f90:466, src/bound.f90:514-598
 *      call solver(n,ng,arrplanp,normfftp,lambdaxyp,ap,bp,cp,cbcpre,['c','c','c'],pp)  src/main.f90:467
 *    The forward x transform evaluates its samples from u, v, w at load time, so the right-hand side is never written to
 *    p and read back (16 B/point less HBM traffic on the two steps); p receives the solution only.  is_bound / have / rhsb
 *    as in cansb200_updt_rhs_b (norm = 1; boundary conditions and c_or_f are the plan's), all three NULL = no wall terms.
 *    Same result as the three separate calls to rounding (the divergence is evaluated by the same expression).  x lengths /
 *    kinds the two-for-one kernels do not serve, and a context with CANSB200_CTX_FUSE_FILLPS = 0, run the three steps one
 *    after the other instead; bit 7 of cansb200_plan_stats()[3] tells which.  Works on one rank and on z slabs (dzfi, u, v, w
 *    are the rank's own slab, as in the reference).  cansb200_solver_fillps is the same through an fftini id. */
int cansb200_solve_fillps(cansb200_plan* plan, void* p, const int n[3], int nhalo, double normfft, const void* lambdaxy,
                          const void* a, const void* b, const void* c, const double dli[3], const void* dzfi, double dti,
                          const void* u, const void* v, const void* w, const int is_bound[6], const int have[3],
                          const double rhsb[6], void* stream);
int cansb200_solver_fillps(cansb200_ctx* ctx, int id, const char bc[6], const char c_or_f[3], void* p, const int n[3], int nhalo,
                           double normfft, const void* lambdaxy, const void* a, const void* b, const void* c, int lambda_order,
                           const double dli[3], const void* dzfi, double dti, const void* u, const void* v, const void* w,
                           const int is_bound[6], const int have[3], const double rhsb[6], void* stream);

/* -- synthetic input: counter-based uniform(-1,1) field indexed by the GLOBAL (i,j,k)
 *    (SURVEY.md 8d); fills the interior of a haloed device array, halo set to 0. */
int cansb200_fill_hash(cansb200_ctx* ctx, void* p, const int n[3], const int lo[3], int nhalo,
                       unsigned long long seed, void* stream);

/* -- diagnostics */
const char* cansb200_last_error(void);
int cansb200_version(void);
/* counters since plan creation: [0] solves, [1] factorisations run, [2] kernels launched (context-wide),
 * [3] bits 0-3 tridiagonal variant in use, bit 4 / 5 pivot cache deduplicated in x / y (after the first solve's check),
 *     bit 6 a later lambdaxy violated the symmetry the deduplicated cache relies on, bit 7 the last cansb200_solve_fillps ran
 *     the fused forward x transform, bits 8-23 y rows per tall tile */
int cansb200_plan_stats(cansb200_plan* plan, unsigned long long stats[4]);
/* per-stage device timing with CUDA events on the solve's stream (the role of the reference's
 * unused timer_tic/toc CUDA-event pool, src/timer.f90:113-216).  Stages of one solve:
 * [0] fft x fwd, [1] fft y fwd, [2] pivot-cache check (+ factorisation on a miss),
 * [3] tridiagonal substitution, [4] fft y bwd, [5] fft x bwd.  ms[] are sums over nsolves. */
int cansb200_set_profiling(cansb200_ctx* ctx, int on);
int cansb200_get_profile(cansb200_ctx* ctx, double ms[8], unsigned long long* nsolves);
/* context switches: what = CANSB200_CTX_FORCE_GENERIC (value 0/1) routes every transform through the
 * generic shared-memory engine instead of the two-for-one register kernels (both are CUDA paths; tests
 * use it to cover the generic engine on lengths the fast path also serves). */
enum {
  CANSB200_CTX_FORCE_GENERIC = 0,
  CANSB200_CTX_X_VARIANT = 1,   /* tuning variant (thread / radix split) of the contiguous transforms: 0..3, -1 = per-kind default (initial state) */
  CANSB200_CTX_Y_VARIANT = 2,   /* same for the strided transforms */
  CANSB200_CTX_CHAIN_COLS = 3,  /* x-window (multiple of 16 columns) of the fft-y -> tridiagonal -> ifft-y chain run window by window on auxiliary streams;
                                   0 = off, -1 = auto (default: two half-width windows, whose kernels overlap each other's tails) */
  CANSB200_CTX_CHAIN_STREAMS = 4, /* auxiliary streams the windows are issued on (1..8) */
  CANSB200_CTX_HOST_CHUNKS = 6,  /* host-memory solves: z-plane chunks whose PCIe copies overlap the x / y transforms (1 = off, default 16) */
  CANSB200_CTX_PIN_HOST = 7,     /* host-memory solves: 1 = page-lock the caller's p on first use (cudaHostRegister; released by
                                    cansb200_finalize) so that a pageable Fortran array copies at the full PCIe rate; default 0 */
  CANSB200_CTX_ZMAJOR = 8,       /* one-GPU solves: the y transforms write / read a z-major copy B[j][k][i] of the field, so that the
                                    tridiagonal stage streams 8 KB-strided rows instead of one row per field plane */
  CANSB200_CTX_DTDMA = 9,        /* several ranks: 1 = the reference's is_poisson_dtdma path: z stays decomposed in the tridiagonal stage
                                    (distributed TDMA), the only exchange is the 2-rows-per-rank reduced system.  Set it BEFORE creating
                                    plans: cansb200_get_extents then reports n_z = (nx, ny, nz_local) as the reference does, i.e. the
                                    caller passes lambdaxy(nx, ny) and its own z slice of a, b, c.  Needs nz >= 6 nranks^2.  Default 0 */
  CANSB200_CTX_DIST_WINDOWS = 10, /* several ranks: x windows of the pipelined exchange (forward y transform of window w + 1, tridiagonal
                                    solve of window w and backward y transform of window w - 1 run concurrently, ordered by
                                    per-window flags); -1 = auto (4 when nx allows), 1 = one window (two whole-field barriers) */
  CANSB200_CTX_DIST_THOMAS_CTAS = 11, /* CTAs of the persistent tridiagonal kernel while it shares the GPU with the y transforms of the
                                    neighbouring windows (one CTA fills an SM); -1 = auto (5/8 of the SMs) */
  CANSB200_CTX_DIST_MODE = 12,    /* several ranks: how the two exchanges travel.  0 = the producing kernels store their rows straight into
                                    the peers' buffers (pack + wire + unpack in one store; the kernel holds its SMs while NVLink drains);
                                    1 = the producing kernels write dense per-destination blocks locally and the copy engines move them
                                    (SMs are free for the HBM-only stages meanwhile, at the price of one more local write + read);
                                    2 = peer stores as 0, but scheduled in two halves: z chunks forward (x transform of chunk c + 1 next to
                                    the y transform of chunk c), x windows backward; -1 = auto */
  CANSB200_CTX_DIST_CHUNKS = 13,  /* copy-engine exchange: z chunks of the forward half (x transform of chunk c + 1 runs while chunk c is on
                                    the wire); -1 = auto */
  CANSB200_CTX_DIST_SPLIT_PAD = 14, /* z-chunked forward half: KB of shared-memory padding per CTA of the y transforms that store to the peers,
                                    so that they leave room on every SM for the x transform of the next chunk; -1 = auto, 0 = none */
  CANSB200_CTX_DTDMA_TILED = 15,  /* distributed TDMA, slab-local elimination: 2 = on chip in the pipelined tridiagonal kernel (24 B/point),
                                    0 = one thread per column sweeping through HBM in the reference's operation order (48 B/point),
                                    -1 / 1 (default) = on chip when the slab has at least 4 rows per thread of a tile (>= 193 rows in FP64) */
  CANSB200_CTX_FUSE_FILLPS = 16,  /* cansb200_solve_fillps: 1 (default) = the forward x transform evaluates fillps (+ updt_rhs_b) at load time,
                                    0 = fillps, updt_rhs_b and the solve run one after the other */
  CANSB200_CTX_AUX_3D = 17,       /* cansb200_fillps / _correc: 1 (default) = kernels with a 3-D launch geometry (no index division per
                                    point), 0 = the flat-index kernels (also the fallback for extents beyond the grid limits) */
  CANSB200_CTX_R2_FLAGS = 5      /* cache hints of the fast transforms: bit 0 = field loads bypass L1 allocation, bit 1 = streaming stores,
                                    bit 2 = force the maximum shared-memory carveout (default: the driver picks, which leaves L1 to the twiddles) */
};
int cansb200_ctx_set(cansb200_ctx* ctx, int what, int value);
/* pencil-sized device buffers the OpenACC host may alias as `work` / `solver_buf_0` / `solver_buf_1` (src/rk.f90:27-29):
 * which = 0, 1, 2.  They are the library's own scratch: any solve on the context overwrites them. */
int cansb200_get_work(cansb200_ctx* ctx, int which, void** ptr, size_t* nelem);

#ifdef __cplusplus
}
#endif
#endif /* CANS_B200_H */
