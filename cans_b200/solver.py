"""Host-side mirror of the reference's solver interface.

Same names, argument order and meaning as the Fortran the C ABI replaces:

    initsolver(...)       /root/reference/src/initsolver.f90:15-83
    fftini / fftend       src/fft.f90:25-245
    solver(...)           src/solver.f90:17-112 == src/solver_gpu.f90:34-276
    solve_helmholtz(...)  src/solve_helmholtz.f90:28-75
    updt_rhs_b            src/bound.f90:514-598

Array convention: a Fortran field p(0:n1+1,0:n2+1,0:n3+1) is a C-ordered array
indexed [k, j, i] of shape (n3+2, n2+2, n1+2) -- a CUDA torch tensor (device
mode, zero copies) or a numpy array (host mode: the library does H2D/D2H).
`cbc[idir][ibound]` with idir 0..2 = x, y, z (the transpose of cbc(0:1,3)).

The arithmetic of `initsolver` (eigenvalues, tridiagonal coefficients) stays
host code exactly as in the reference, where it is Fortran run once at start-up.
Everything per-solve runs in the CUDA library; nothing here falls back to a
CPU implementation.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, Options, check, d3, i3, lib

try:  # torch is plumbing only: device memory + streams
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _ptr(x) -> int:
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    raise TypeError(f"unsupported array type {type(x)}")


def _stream_ptr(stream=None) -> int:
    if stream is not None:
        return int(stream.cuda_stream) if hasattr(stream, "cuda_stream") else int(stream)
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_stream().cuda_stream)
    return 0


_current_ctx = None   # the most recently created context: what the Fortran shim keeps as a module variable


class Context:
    """cansb200_ctx: replaces initmpi's decomposition setup + workspaces.  One rank: dims = [1, 1];
    several ranks (one per GPU of a box): dims = [1, nranks], z slabs, followed by `connect()`."""

    def __init__(self, ng: Sequence[int], is_fp32: bool = False, dims=None, ipencil_axis: int = 1, rank: int = 0,
                 nranks: int = 1, nccl_id: Optional[bytes] = None):
        self.ng = [int(v) for v in ng]
        self.is_fp32 = bool(is_fp32)
        self.dtype = np.float32 if is_fp32 else np.float64
        self.rank, self.nranks = int(rank), int(nranks)
        dims = (1, self.nranks) if dims is None else dims
        self._h = C.c_void_p()
        idp = C.cast(C.c_char_p(nccl_id), C.c_void_p) if nccl_id else None
        check(lib.cansb200_init(C.byref(self._h), i3(self.ng), i3(dims), ipencil_axis, rank, nranks, idp, int(is_fp32)),
              "cansb200_init")
        n, lo, nz, loz = i3([0] * 3), i3([0] * 3), i3([0] * 3), i3([0] * 3)
        check(lib.cansb200_get_extents(self._h, n, lo, nz, loz), "get_extents")
        self.n, self.lo, self.n_z, self.lo_z = list(n), list(lo), list(nz), list(loz)
        global _current_ctx
        _current_ctx = self

    def get_work(self, which: int):
        """-> (device pointer, number of elements) of the buffer an OpenACC host aliases as `work` (0), `solver_buf_0` (1),
        `solver_buf_1` (2) (src/rk.f90:27-29)."""
        ptr, nel = C.c_void_p(), C.c_size_t()
        check(lib.cansb200_get_work(self._h, int(which), C.byref(ptr), C.byref(nel)), "get_work")
        return ptr.value, nel.value

    def connect(self, group=None):
        """Rendezvous of the ranks of one box: gather every rank's IPC blob (torch.distributed plays the role
        MPI_Allgather has in a Fortran host) and map the peers' exchange regions."""
        if self.nranks < 2:
            return
        import torch.distributed as dist
        nb = lib.cansb200_dist_blob_size()
        mine = (C.c_ubyte * nb)()
        check(lib.cansb200_dist_export(self._h, mine), "dist_export")
        blobs = [None] * self.nranks
        dist.all_gather_object(blobs, bytes(mine), group=group)
        allb = b"".join(blobs)
        buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
        check(lib.cansb200_dist_connect(self._h, buf), "dist_connect")
        dist.barrier(group=group)

    @staticmethod
    def connect_local(ctxs):
        """All ranks in THIS process on THIS device (ctxs[r] = context of rank r): wires the exchange regions directly
        (`cansb200_dist_connect_local`).  Each rank's solve must be enqueued on its own stream before any host sync."""
        arr = (C.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
        check(lib.cansb200_dist_connect_local(arr, len(ctxs)), "dist_connect_local")

    def is_bound(self):
        """`is_bound(0:1,3)` of initmpi (src/initmpi.f90:193-250) as [idir][ibound]: does this rank own the physical
        boundary?  z slabs: x and y always, z only on the first / last rank."""
        return [[True, True], [True, True],
                [self.lo[2] == 1, self.lo[2] + self.n[2] - 1 == self.ng[2]]]

    def hi(self):
        return [self.lo[d] + self.n[d] - 1 for d in range(3)]

    def set_dist_windows(self, windows: int = -1, thomas_ctas: int = -1):
        """Several ranks: x windows of the pipelined exchange (-1 = auto, 1 = two whole-field barriers) and the CTAs the
        tridiagonal kernel may take while it shares the GPU with the y transforms."""
        check(lib.cansb200_ctx_set(self._h, 10, int(windows)), "ctx_set")
        check(lib.cansb200_ctx_set(self._h, 11, int(thomas_ctas)), "ctx_set")

    def set_dist_mode(self, mode: int = -1, chunks: int = -1, split_pad_kb: int = -1):
        """Several ranks: 0 = the producing kernels store straight into the peers' buffers, 1 = dense local send blocks moved
        by the copy engines (z chunks forward, x windows back), -1 = auto."""
        check(lib.cansb200_ctx_set(self._h, 12, int(mode)), "ctx_set")
        check(lib.cansb200_ctx_set(self._h, 13, int(chunks)), "ctx_set")
        check(lib.cansb200_ctx_set(self._h, 14, int(split_pad_kb)), "ctx_set")

    def dist_status(self) -> int:
        st = C.c_int()
        check(lib.cansb200_dist_status(self._h, C.byref(st)), "dist_status")
        return int(st.value)

    @property
    def handle(self):
        return self._h

    def force_generic_engine(self, on: bool = True):
        """Route every transform through the generic shared-memory engine (tests)."""
        check(lib.cansb200_ctx_set(self._h, 0, int(on)), "ctx_set")

    def set_variant(self, x: int = -1, y: int = -1):
        """Tuning variants of the fast transforms (thread / radix split), see r2r2_inst.cuh."""
        check(lib.cansb200_ctx_set(self._h, 1, int(x)), "ctx_set")
        check(lib.cansb200_ctx_set(self._h, 2, int(y)), "ctx_set")

    def set_r2_flags(self, flags: int = 0):
        """Cache hints of the fast transforms (CANSB200_CTX_R2_FLAGS): 1 = field loads bypass L1
        allocation, 2 = streaming stores, 4 = force the maximum shared-memory carveout."""
        check(lib.cansb200_ctx_set(self._h, 5, int(flags)), "ctx_set")

    def set_host_chunks(self, n: int = 16):
        """Host-memory solves: number of z-plane chunks whose copies overlap the transforms (1 = one copy each way)."""
        check(lib.cansb200_ctx_set(self._h, 6, int(n)), "ctx_set")

    def set_dtdma(self, on: bool = True):
        """Several ranks: the reference's `is_poisson_dtdma` path (distributed TDMA, z stays decomposed).  Call it before
        `initsolver`: `n_z` / `lo_z` become those of the slab, as in the reference (src/solver.f90:43-48)."""
        check(lib.cansb200_ctx_set(self._h, 9, int(on)), "ctx_set")
        n, lo, nz, loz = i3([0] * 3), i3([0] * 3), i3([0] * 3), i3([0] * 3)
        check(lib.cansb200_get_extents(self._h, n, lo, nz, loz), "get_extents")
        self.n, self.lo, self.n_z, self.lo_z = list(n), list(lo), list(nz), list(loz)

    def set_dtdma_tiled(self, on=None):
        """Distributed TDMA, slab-local elimination: True = always on chip (pipelined kernel), False = per-column sweeps in the
        reference's operation order, None = automatic (on chip for slabs of at least 193 rows in FP64)."""
        check(lib.cansb200_ctx_set(self._h, 15, 1 if on is None else (2 if on else 0)), "ctx_set")

    def set_fuse_fillps(self, on: bool = True):
        """`solver_fillps`: evaluate fillps (+ updt_rhs_b) inside the forward x transform (default) or run the three steps
        one after the other."""
        check(lib.cansb200_ctx_set(self._h, 16, int(on)), "ctx_set")

    def set_aux_3d(self, on: bool = True):
        """`fillps` / `correc`: kernels with the 3-D launch geometry (default) or the flat-index ones."""
        check(lib.cansb200_ctx_set(self._h, 17, int(on)), "ctx_set")

    def set_zmajor(self, on: bool = True):
        """One-GPU solves: z-major intermediate between the y transforms and the tridiagonal stage."""
        check(lib.cansb200_ctx_set(self._h, 8, int(on)), "ctx_set")

    def set_pin_host(self, on: bool = True):
        """Host-memory solves: page-lock the caller's array on first use (cudaHostRegister)."""
        check(lib.cansb200_ctx_set(self._h, 7, int(on)), "ctx_set")

    def set_chain(self, cols: int = 0, streams: int = 2):
        """L2-resident fft-y -> tridiagonal -> ifft-y chain over x windows of `cols` columns (0 = off)."""
        check(lib.cansb200_ctx_set(self._h, 3, int(cols)), "ctx_set")
        check(lib.cansb200_ctx_set(self._h, 4, int(streams)), "ctx_set")

    def set_profiling(self, on: bool):
        check(lib.cansb200_set_profiling(self._h, int(on)), "set_profiling")

    def get_profile(self):
        """-> (dict stage -> total ms, number of profiled solves)"""
        ms = (C.c_double * 8)()
        ns = C.c_ulonglong()
        check(lib.cansb200_get_profile(self._h, ms, C.byref(ns)), "get_profile")
        names = ["fft_x_fwd", "fft_y_fwd", "pivot_cache", "thomas", "fft_y_bwd", "fft_x_bwd"]
        return {k: ms[i] for i, k in enumerate(names)}, int(ns.value)

    def close(self):
        if self._h:
            lib.cansb200_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """What `arrplan(2,2)` is in the reference: the four r2r plans of one solved variable
    (the handle a Fortran host would keep in arrplan(1,1) as type(C_PTR))."""

    def __init__(self, ctx: Context, cbc, c_or_f, **options):
        self.ctx = ctx
        self.cbc = [list(b) for b in cbc]
        self.c_or_f = list(c_or_f)
        bc6 = "".join(self.cbc[d][i] for d in range(3) for i in range(2)).encode()
        cf3 = "".join(self.c_or_f).encode()
        self._h = C.c_void_p()
        nf = C.c_double()
        opt = Options(**options)
        check(lib.cansb200_plan_create(ctx.handle, C.byref(self._h), bc6, cf3, C.byref(opt), C.byref(nf)), "plan_create")
        self.normfft = ctx.dtype(nf.value)

    @property
    def handle(self):
        return self._h

    @property
    def id(self) -> int:
        """integer name of the plan: what arrplan(1,1) holds on the reference's CUDA build (`integer, dimension(2,2)`)"""
        return int(lib.cansb200_plan_id(self._h))

    @staticmethod
    def handle_from_id(i: int):
        return lib.cansb200_plan_from_id(int(i))

    def stats(self):
        s = (C.c_ulonglong * 4)()
        check(lib.cansb200_plan_stats(self._h, s), "plan_stats")
        return {"solves": s[0], "factorisations": s[1], "launches": s[2], "thomas_variant": s[3] & 15,
                "pivot_dedup_x": (s[3] >> 4) & 1, "pivot_dedup_y": (s[3] >> 5) & 1,
                # set if a lambdaxy WITHOUT the mirror symmetry reached a plan whose cache is deduplicated after its first solve
                # (the first solve checks and falls back; later solves only record the violation here)
                "pivot_dedup_violation": (s[3] >> 6) & 1, "fillps_fused": (s[3] >> 7) & 1, "tall_tile_rows": (s[3] >> 8) & 0xFFFF}

    def destroy(self):
        if self._h:
            lib.cansb200_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# initsolver and its pieces (host arithmetic, as in the reference)
# ---------------------------------------------------------------------------
def eigenvalues(n, bc, c_or_f, dtype=np.float64, openacc_order=False):
    """src/initsolver.f90:85-144.  `openacc_order`: the `#if defined(_OPENACC)` permutation of a periodic direction
    (:98-117, `iswap`): (r0, r[n/2], r1, i1, ...) instead of the CPU build's halfcomplex order; a plan created with
    `lambda_order=1` takes lambdaxy in that order."""
    pi = np.arccos(dtype(-1.0))
    l = np.arange(1, n + 1)
    key = bc[0] + bc[1]
    two, one = dtype(2.0), dtype(1.0)
    if key == "PP":
        th = (2 * (l - 1)).astype(dtype) * pi / dtype(n)
    elif key == "NN":
        th = (l - 1).astype(dtype) * pi / dtype(n)
    elif key == "DD":
        th = l.astype(dtype) * pi / dtype(n)
    else:
        th = (2 * l - 1).astype(dtype) * pi / dtype(2.0 * n)
    lam = (-two * (one - np.cos(th))).astype(dtype)
    if key == "DD" and c_or_f == "f":
        lam[n - 1] = 0.0
    if key == "PP" and openacc_order:
        nh = (n + 1) // 2
        iswap = np.zeros(n + 1, dtype=np.int64)   # 1-based, exactly the reference's loop
        iswap[1] = 1
        if n >= 2:
            iswap[2] = nh + (1 - n % 2)
        for ll in range(2, n):
            if ll <= nh:
                iswap[2 * ll - 1] = ll
            else:
                iswap[n - 2 * (ll - (nh + 1)) - n % 2] = ll + 1
        lam = lam[iswap[1:] - 1]
    return lam


def tridmatrix(bc, n, dzci, dzfi, c_or_f, dtype=np.float64):
    """src/initsolver.f90:146-187; dzci / dzfi indexed 0..n+1."""
    k = np.arange(1, n + 1)
    if c_or_f == "c":
        a = (dzfi[k] * dzci[k - 1]).astype(dtype)
        c = (dzfi[k] * dzci[k]).astype(dtype)
    else:
        a = (dzfi[k] * dzci[k]).astype(dtype)
        c = (dzfi[k + 1] * dzci[k]).astype(dtype)
    b = -(a + c)
    f = [{"P": 0.0, "D": -1.0, "N": 1.0}[bc[i]] for i in (0, 1)]
    if c_or_f == "c" or bc[0] == "N":
        b[0] = b[0] + dtype(f[0]) * a[0]
    if c_or_f == "c" or bc[1] == "N":
        b[n - 1] = b[n - 1] + dtype(f[1]) * c[n - 1]
    return a, b, c


def bc_rhs(cbc, bc, dlc, dlf, c_or_f):
    """src/initsolver.f90:189-232 -> [lower, upper] plane values."""
    out = []
    for ib in (0, 1):
        sgn = 1.0 if ib == 0 else -1.0
        if cbc[ib] == "P":
            fac = 0.0
        elif cbc[ib] == "D":
            fac = -2.0 * bc[ib] if c_or_f == "c" else -bc[ib]
        else:
            fac = sgn * (dlc[ib] if c_or_f == "c" else dlf[ib]) * bc[ib]
        out.append(fac / dlc[ib] / dlf[ib])
    return out


def find_fft(bc, c_or_f):
    """src/fft.f90:260-313 -> (kind_fwd, kind_bwd, (norm1, norm2)) with FFTW kind numbers."""
    key = bc[0] + bc[1]
    if key == "PP":
        return 0, 1, (1.0, 0.0)
    if c_or_f == "c":
        return {"NN": (5, 4, (2.0, 0.0)), "DD": (9, 8, (2.0, 0.0)), "ND": (6, 6, (2.0, 0.0)), "DN": (10, 10, (2.0, 0.0))}[key]
    return {"NN": (3, 3, (2.0, -1.0)), "DD": (7, 7, (2.0, 1.0)), "ND": (5, 4, (2.0, 0.0)), "DN": (8, 9, (2.0, 0.0))}[key]


def fftini(ctx: Context, cbc, c_or_f, **options):
    """src/fft.f90:25-209 -> (arrplan, normfft).  `cbc` needs all three directions because the
    plan also fixes the z-solve variant (periodic closure, face-centred Dirichlet exclusion)."""
    plan = Plan(ctx, cbc, c_or_f, **options)
    return plan, plan.normfft


def fftend(arrplan: Plan):
    """src/fft.f90:211-245."""
    arrplan.destroy()


@dataclass
class SolverData:
    lambdaxy: Any
    a: Any
    b: Any
    c: Any
    arrplan: Plan
    normfft: Any
    rhsbx: List[float]
    rhsby: List[float]
    rhsbz: List[float]
    host: dict = field(default_factory=dict)


def initsolver(ctx: Context, ng, dli, dzci_g, dzfi_g, cbc, bc, c_or_f, device=None, openacc_order=False, **options) -> SolverData:
    """src/initsolver.f90:15-83.  Returns lambdaxy[j,i], a, b, c (on `device` if given, else numpy),
    arrplan, normfft and the boundary RHS values.  `openacc_order`: behave like the _OPENACC build of initsolver
    (packed eigenvalue order in periodic directions) and create the plan with `lambda_order=1`."""
    dt = ctx.dtype
    if openacc_order:
        options["lambda_order"] = 1
    dli = [dt(v) for v in dli]
    dzci_g = np.asarray(dzci_g, dtype=dt)
    dzfi_g = np.asarray(dzfi_g, dtype=dt)
    lx = eigenvalues(ng[0], cbc[0], c_or_f[0], dt, openacc_order) * dli[0] ** 2
    ly = eigenvalues(ng[1], cbc[1], c_or_f[1], dt, openacc_order) * dli[1] ** 2
    # lambdaxy(lo_z(1):hi_z(1), lo_z(2):hi_z(2)): the z-pencil slice of this rank (src/initsolver.f90:54-58)
    y0 = ctx.lo_z[1] - 1
    ly = ly[y0:y0 + ctx.n_z[1]]
    lambdaxy = np.ascontiguousarray((lx[None, :] + ly[:, None]).astype(dt))
    a, b, c = tridmatrix(cbc[2], ng[2], dzci_g, dzfi_g, c_or_f[2], dt)
    # a, b, c(lo_z(3):hi_z(3)) (src/initsolver.f90:60-65): the whole z range except in distributed-TDMA mode
    k0 = ctx.lo_z[2] - 1
    a, b, c = (np.ascontiguousarray(v[k0:k0 + ctx.n_z[2]]) for v in (a, b, c))
    dl = [dt(1.0) / v for v in dli]
    dzc_g, dzf_g = dt(1.0) / dzci_g, dt(1.0) / dzfi_g
    n3 = ng[2]
    rhsbx = bc_rhs(cbc[0], bc[0], [dl[0]] * 2, [dl[0]] * 2, c_or_f[0])
    rhsby = bc_rhs(cbc[1], bc[1], [dl[1]] * 2, [dl[1]] * 2, c_or_f[1])
    if c_or_f[2] == "c":
        rhsbz = bc_rhs(cbc[2], bc[2], [dzc_g[0], dzc_g[n3]], [dzf_g[1], dzf_g[n3]], "c")
    else:
        rhsbz = bc_rhs(cbc[2], bc[2], [dzc_g[1], dzc_g[n3 - 1]], [dzf_g[1], dzf_g[n3]], "f")
    arrplan, normfft = fftini(ctx, cbc, c_or_f, **options)
    host = dict(lambdaxy=lambdaxy, a=a, b=b, c=c)
    if device is not None:
        lam_d = torch.from_numpy(lambdaxy).to(device)
        a_d, b_d, c_d = (torch.from_numpy(v).to(device) for v in (a, b, c))
        return SolverData(lam_d, a_d, b_d, c_d, arrplan, normfft, rhsbx, rhsby, rhsbz, host)
    return SolverData(lambdaxy, a, b, c, arrplan, normfft, rhsbx, rhsby, rhsbz, host)


# ---------------------------------------------------------------------------
# solver / solve_helmholtz
# ---------------------------------------------------------------------------
def solver(n, ng, arrplan: Plan, normfft, lambdaxy, a, b, c, bc, c_or_f, p, stream=None):
    """src/solver.f90:17-112.  Solves in place on the interior of the haloed `p`.
    `bc` / `c_or_f` are accepted for signature parity; the plan already fixed them."""
    if list(c_or_f) != arrplan.c_or_f or [list(x) for x in bc] != arrplan.cbc:
        raise ValueError("solver: bc / c_or_f differ from the ones the plan (arrplan) was created with")
    dev = _is_torch(p) and p.is_cuda
    for name, arr in (("lambdaxy", lambdaxy), ("a", a), ("b", b), ("c", c)):
        if (_is_torch(arr) and arr.is_cuda) != dev:
            raise ValueError(f"solver: {name} must live where p lives (all device or all host)")
    want = (n[2] + 2, n[1] + 2, n[0] + 2)
    if tuple(p.shape) != want:
        raise ValueError(f"solver: p has shape {tuple(p.shape)}, expected {want}")
    check(lib.cansb200_solve(arrplan.handle, _ptr(p), i3(n), 1, float(normfft), _ptr(lambdaxy), _ptr(a), _ptr(b), _ptr(c),
                             MEM_DEVICE if dev else MEM_HOST, _stream_ptr(stream)), "cansb200_solve")
    return p


def solver_fillps(n, ng, arrplan: Plan, normfft, lambdaxy, a, b, c, bc, c_or_f, dli, dzfi, dti, u, v, w, p,
                  is_bound=None, rhsbx=None, rhsby=None, rhsbz=None, stream=None):
    """`fillps` + `updt_rhs_b` + `solver` of the pressure-correction step (src/main.f90:465-467) as ONE call on device
    arrays: the forward x transform evaluates the right-hand side from u, v, w at load time (`cansb200_solve_fillps`), p only
    receives the solution.  The first ten arguments are `solver`'s, `dli, dzfi, dti, u, v, w` are `fillps`'s, `is_bound,
    rhsbx, rhsby, rhsbz` are `updt_rhs_b`'s ([lower, upper] wall values or None)."""
    if list(c_or_f) != arrplan.c_or_f or [list(x) for x in bc] != arrplan.cbc:
        raise ValueError("solver_fillps: bc / c_or_f differ from the ones the plan (arrplan) was created with")
    for name, arr in (("p", p), ("u", u), ("v", v), ("w", w), ("dzfi", dzfi), ("lambdaxy", lambdaxy), ("a", a), ("b", b), ("c", c)):
        if not (_is_torch(arr) and arr.is_cuda):
            raise ValueError(f"solver_fillps: {name} must be a CUDA tensor (device arrays only)")
    want = (n[2] + 2, n[1] + 2, n[0] + 2)
    for name, arr in (("p", p), ("u", u), ("v", v), ("w", w)):
        if tuple(arr.shape) != want:
            raise ValueError(f"solver_fillps: {name} has shape {tuple(arr.shape)}, expected {want}")
    if tuple(dzfi.shape) != (n[2] + 2,):
        raise ValueError("solver_fillps: dzfi must have n(3)+2 entries (0:n3+1)")
    walls = any(r is not None for r in (rhsbx, rhsby, rhsbz))
    isb = have = vals = None
    if walls:
        if is_bound is None:
            raise ValueError("solver_fillps: is_bound is needed with rhsbx / rhsby / rhsbz")
        rh = [r if r is not None else [0.0, 0.0] for r in (rhsbx, rhsby, rhsbz)]
        have = (C.c_int * 3)(*[int(r is not None) for r in (rhsbx, rhsby, rhsbz)])
        isb = (C.c_int * 6)(*[int(bool(is_bound[d][sd])) for d in range(3) for sd in range(2)])
        vals = (C.c_double * 6)(*[float(rh[d][sd]) for d in range(3) for sd in range(2)])
    check(lib.cansb200_solve_fillps(arrplan.handle, _ptr(p), i3(n), 1, float(normfft), _ptr(lambdaxy), _ptr(a), _ptr(b), _ptr(c),
                                    d3(dli), _ptr(dzfi), float(dti), _ptr(u), _ptr(v), _ptr(w), isb, have, vals,
                                    _stream_ptr(stream)), "cansb200_solve_fillps")
    return p


def solver_gaussel_z(n, ng, hi, a, b, c, bcz, c_or_f, norm, p, arrplan: Plan = None, stream=None, ctx: Context = None):
    """src/solver.f90:547-616 (`solver_gaussel_z`, the implicit-z-diffusion solve of `is_impdiff_1d`):
    lambda-less tridiagonal solve in z on the interior of the haloed `p`, in place.  The first ten arguments are the
    reference's (`call solver_gaussel_z(n,ng,hi,a,bb,c,cbc(:,3),c_or_f,alphai,p)`, src/solve_helmholtz.f90:73): no plan
    is needed -- the context (the most recently created one unless `ctx=` is given, like the module variable of the Fortran
    shim) keeps one z-only plan per (bcz, c_or_f(3)) (`cansb200_solve_z_bc`).  Passing `arrplan=` uses that plan instead."""
    dev = _is_torch(p) and p.is_cuda
    for name, arr in (("a", a), ("b", b), ("c", c)):
        if (_is_torch(arr) and arr.is_cuda) != dev:
            raise ValueError(f"solver_gaussel_z: {name} must live where p lives (all device or all host)")
    want = (n[2] + 2, n[1] + 2, n[0] + 2)
    if tuple(p.shape) != want:
        raise ValueError(f"solver_gaussel_z: p has shape {tuple(p.shape)}, expected {want}")
    mk = MEM_DEVICE if dev else MEM_HOST
    if arrplan is not None:
        if list(bcz) != arrplan.cbc[2] or c_or_f[2] != arrplan.c_or_f[2]:
            raise ValueError("solver_gaussel_z: bcz / c_or_f(3) differ from the ones the plan was created with")
        check(lib.cansb200_solve_z(arrplan.handle, _ptr(p), i3(n), 1, float(norm), _ptr(a), _ptr(b), _ptr(c), mk, _stream_ptr(stream)),
              "cansb200_solve_z")
        return p
    ctx = ctx or _current_ctx
    if ctx is None:
        raise ValueError("solver_gaussel_z: no context exists yet")
    check(lib.cansb200_solve_z_bc(ctx.handle, (bcz[0] + bcz[1]).encode(), c_or_f[2].encode(), _ptr(p), i3(n), 1, float(norm),
                                  _ptr(a), _ptr(b), _ptr(c), mk, _stream_ptr(stream)), "cansb200_solve_z_bc")
    return p


def updt_rhs_b(c_or_f, cbc, n, is_bound, rhsbx, rhsby, rhsbz, p, alpha=None, ctx: Context = None, stream=None):
    """src/bound.f90:514-598: adds the wall contributions to the first / last interior planes of the ranks that own
    the wall (`is_bound[idir][ibound]`, as `Context.is_bound()` returns it)."""
    norm = 1.0 if alpha is None else alpha
    if _is_torch(p) and p.is_cuda:
        # device arrays: one CUDA kernel per direction (cansb200_updt_rhs_b), stream ordered like the solve that follows
        ctx = ctx or _current_ctx
        rh = [v if v is not None else [0.0, 0.0] for v in (rhsbx, rhsby, rhsbz)]
        have = (C.c_int * 3)(*[int(v is not None) for v in (rhsbx, rhsby, rhsbz)])
        isb = (C.c_int * 6)(*[int(bool(is_bound[d][sd])) for d in range(3) for sd in range(2)])
        vals = (C.c_double * 6)(*[float(rh[d][sd]) for d in range(3) for sd in range(2)])
        bc6 = "".join(cbc[d][i] for d in range(3) for i in range(2)).encode()
        check(lib.cansb200_updt_rhs_b(ctx.handle, "".join(c_or_f).encode(), bc6, i3(n), isb, have, vals, float(norm), _ptr(p),
                                      _stream_ptr(stream)), "cansb200_updt_rhs_b")
        return
    q = [1 if (c_or_f[d] == "f" and cbc[d][1] == "D") else 0 for d in range(3)]
    n1, n2, n3 = n
    K, J, I = slice(1, n3 + 1), slice(1, n2 + 1), slice(1, n1 + 1)
    if rhsbx is not None:
        if is_bound[0][0]:
            p[K, J, 1] += rhsbx[0] * norm
        if is_bound[0][1]:
            p[K, J, n1 - q[0]] += rhsbx[1] * norm
    if rhsby is not None:
        if is_bound[1][0]:
            p[K, 1, I] += rhsby[0] * norm
        if is_bound[1][1]:
            p[K, n2 - q[1], I] += rhsby[1] * norm
    if rhsbz is not None:
        if is_bound[2][0]:
            p[1, J, I] += rhsbz[0] * norm
        if is_bound[2][1]:
            p[n3 - q[2], J, I] += rhsbz[1] * norm


def solve_helmholtz(n, ng, hi, arrplan, normfft, alpha, lambdaxy, a, b, c, rhsbx, rhsby, rhsbz, is_bound, cbc, c_or_f, p,
                    stream=None, is_impdiff_1d=False):
    """src/solve_helmholtz.f90:28-75 (same argument order): p/alpha + lap(p) = rhs.  `is_impdiff_1d` is a module
    parameter in the reference (`mod_param`); here a keyword."""
    updt_rhs_b(c_or_f, cbc, n, is_bound, rhsbx, rhsby, rhsbz, p, alpha, ctx=arrplan.ctx if arrplan is not None else None, stream=stream)
    ty = p.dtype.type if isinstance(p, np.ndarray) else (np.float32 if p.dtype == torch.float32 else np.float64)
    alphai = ty(1.0) / ty(alpha)
    if _is_torch(b):
        # `bb(k) = b(k) + alphai` runs on the solve's stream (the reference does it on the same OpenACC queue): on any other
        # stream the solve's kernels could read bb before it is written
        if stream is not None and b.is_cuda:
            with torch.cuda.stream(stream if hasattr(stream, "cuda_stream") else torch.cuda.ExternalStream(int(stream))):
                bb = b + float(alphai)
        else:
            bb = b + float(alphai)
    else:
        bb = b + alphai
    if is_impdiff_1d:
        return solver_gaussel_z(n, ng, hi, a, bb, c, cbc[2], c_or_f, alphai, p, arrplan=arrplan, stream=stream)
    return solver(n, ng, arrplan, ty(normfft) * alphai, lambdaxy, a, bb, c, cbc, c_or_f, p, stream)


# ---------------------------------------------------------------------------
# stage-level calls + the steps either side of the path (device tensors)
# ---------------------------------------------------------------------------
def r2r(ctx: Context, kind: int, n_transform: int, axis: int, arr, stream=None):
    """One `call fft(plan, arr)` (src/fft.f90:247-258) on a haloless device array [k,j,i]."""
    nz, ny, nx = arr.shape
    check(lib.cansb200_r2r(ctx.handle, kind, n_transform, axis, _ptr(arr), i3([nx, ny, nz]), _stream_ptr(stream)), "cansb200_r2r")
    return arr


def gaussel(arrplan: Plan, n_rows, a, b, c, is_periodic, norm, pz, lambdaxy, stream=None):
    """`call gaussel(nx,ny,n,0,a,b,c,is_periodic,norm,p,lambdaxy)` (src/solver.f90:114-307)."""
    nz, ny, nx = pz.shape
    check(lib.cansb200_gaussel(arrplan.handle, _ptr(pz), i3([nx, ny, nz]), int(n_rows), int(is_periodic), float(norm),
                               _ptr(lambdaxy), _ptr(a), _ptr(b), _ptr(c), _stream_ptr(stream)), "cansb200_gaussel")
    return pz


def gaussel_dtdma(arrplan: Plan, starts, n_rows, a, b, c, is_periodic, norm, pz, lambdaxy=None, stream=None):
    """`call gaussel_dtdma(nx,ny,n,0,a,b,c,is_periodic,norm,p,lambdaxy)` (src/solver.f90:309-517) with the z slabs of
    len(starts) - 1 ranks on one GPU; `pz[k, j, i]` holds the rows of the global system."""
    nz, ny, nx = pz.shape
    st = (C.c_int * len(starts))(*[int(v) for v in starts])
    check(lib.cansb200_gaussel_dtdma(arrplan.handle, _ptr(pz), i3([nx, ny, nz]), int(n_rows), len(starts) - 1, st,
                                     int(is_periodic), float(norm), _ptr(lambdaxy) if lambdaxy is not None else None,
                                     _ptr(a), _ptr(b), _ptr(c), _stream_ptr(stream)), "cansb200_gaussel_dtdma")
    return pz


def fillps(ctx, n, dli, dzfi, dti, u, v, w, p, stream=None):
    """src/fillps.f90:13-51."""
    check(lib.cansb200_fillps(ctx.handle, i3(n), d3(dli), _ptr(dzfi), float(dti), _ptr(u), _ptr(v), _ptr(w), _ptr(p),
                              _stream_ptr(stream)), "cansb200_fillps")


def correc(ctx, n, dli, dzci, dt, p, u, v, w, stream=None):
    """src/correc.f90:13-60."""
    check(lib.cansb200_correc(ctx.handle, i3(n), d3(dli), _ptr(dzci), float(dt), _ptr(p), _ptr(u), _ptr(v), _ptr(w),
                              _stream_ptr(stream)), "cansb200_correc")


def chkdiv(ctx, n, l, dli, dzfi, u, v, w, stream=None):
    """src/chkdiv.f90:15-54 -> (divtot, divmax)."""
    tot, mx = C.c_double(), C.c_double()
    check(lib.cansb200_chkdiv(ctx.handle, i3(n), d3(dli), _ptr(dzfi), _ptr(u), _ptr(v), _ptr(w), C.byref(tot), C.byref(mx),
                              _stream_ptr(stream)), "cansb200_chkdiv")
    return tot.value / float(l[0] * l[1] * l[2]), mx.value


def fill_hash(ctx, p, n, lo, nhalo, seed, stream=None):
    """Counter-based synthetic field written on the device."""
    check(lib.cansb200_fill_hash(ctx.handle, _ptr(p), i3(n), i3(lo), int(nhalo), int(seed), _stream_ptr(stream)),
          "cansb200_fill_hash")
    return p
