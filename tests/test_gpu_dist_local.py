"""The multi-GPU (z-slab) solve on ONE GPU: P ranks = P contexts of this process, wired with
`cansb200_dist_connect_local`.  Same kernels (SPLIT y transforms with peer row tables, tridiagonal solve with
`out_rows`, device-side flag kernels), same pipeline over x windows as on a multi-GPU box -- only the peers'
regions are plain device pointers instead of CUDA-IPC mappings.  This is what a single-GPU `-m gpu` run can
prove of row a7 (transposes) and e (multi-GPU) of SURVEY.md section 8; tests/test_gpu_dist.py runs the real thing
on 2/4/8 GPUs and bench.py --gpus N checks the distributed result against a single-GPU solve.

Reference being replaced: the transposes around `gaussel` in src/solver.f90:62-107 / src/solver_gpu.f90:99-258."""
import importlib

import numpy as np
import pytest

import cases
from oracle import cans_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    return importlib.import_module("cans_b200")


def _run_virtual_ranks(cb, name, P, windows, mode=0, chunks=-1):
    from cans_b200.decomp import SlabDecomp
    ng, l, cbc, cf, gr, dt, helm = cases.DIST_CASES[name]
    if P > ng[1] or P > ng[2]:
        pytest.skip("more ranks than planes")
    dev = torch.device("cuda:0")
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
    p = cases.make_rhs(cs)
    ref = p.copy()
    if helm:
        O.solve_helmholtz(ng, ng, cs["arrplan"], cs["normfft"], cases.ALPHA, cs["lambdaxy"], cs["a"], cs["b"], cs["c"],
                          None, None, None, cbc, cf, ref)
    else:
        O.solver(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cbc, cf, ref)
    ctxs = [cb.Context(ng, is_fp32=dt == np.float32, rank=r, nranks=P) for r in range(P)]
    for r, c in enumerate(ctxs):
        dec = SlabDecomp(ng, P, r)
        assert c.n == dec.n and c.lo == dec.lo and c.n_z == dec.n_z and c.lo_z == dec.lo_z
        c.set_dist_windows(windows)
        c.set_dist_mode(mode, chunks)
    cb.Context.connect_local(ctxs)
    sds = [cb.initsolver(c, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev) for c in ctxs]
    streams = [torch.cuda.Stream() for _ in range(P)]
    slabs_h, slabs_d = [], []
    for r in range(P):
        z0, z1 = SlabDecomp(ng, P, r).z_range()
        h = np.zeros((z1 - z0 + 2, ng[1] + 2, ng[0] + 2), dtype=dt)
        h[1:-1] = p[1 + z0:1 + z1]
        slabs_h.append(h)
        slabs_d.append(torch.from_numpy(h).to(dev))
    torch.cuda.synchronize()
    for rep in range(3):   # repeated solves exercise buffer and flag reuse
        for r in range(P):   # every rank's solve is enqueued before anything synchronises
            with torch.cuda.stream(streams[r]):
                slabs_d[r].copy_(torch.from_numpy(slabs_h[r]), non_blocking=False)
        torch.cuda.synchronize()
        for r in range(P):
            c, sd = ctxs[r], sds[r]
            if helm:
                cb.solve_helmholtz(c.n, ng, c.hi(), sd.arrplan, sd.normfft, cases.ALPHA, sd.lambdaxy, sd.a, sd.b, sd.c, None, None,
                                   None, c.is_bound(), cbc, cf, slabs_d[r], stream=streams[r])
            else:
                cb.solver(c.n, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, slabs_d[r], stream=streams[r])
    torch.cuda.synchronize()
    for c in ctxs:
        assert c.dist_status() == 0, "a device-side wait timed out"
    got = [t.cpu().numpy() for t in slabs_d]
    full = np.concatenate([g[1:-1, 1:-1, 1:-1] for g in got], axis=0)
    err = cases.parity_error(cs, full, ref[1:-1, 1:-1, 1:-1], helm)
    tol = 1e-12 if dt == np.float64 else 1e-5
    assert err < tol, f"{name} P={P} windows={windows}: rel L2 {err:.3e}"
    for r in range(P):
        halo = np.ones(got[r].shape, bool)
        halo[1:-1, 1:-1, 1:-1] = False
        assert np.array_equal(got[r][halo], slabs_h[r][halo]), "halo cells were modified"
    # z-only solve (solver_gaussel_z) over the same exchange
    alphai = dt(1.0) / dt(cases.ALPHA)
    bb = (cs["b"] + alphai).astype(dt)
    refz = p.copy()
    O.solver_gaussel_z(ng, ng, ng, cs["a"], bb, cs["c"], cbc[2], cf, alphai, refz)
    bbd = torch.from_numpy(bb).to(dev)
    for r in range(P):
        slabs_d[r].copy_(torch.from_numpy(slabs_h[r]))
    torch.cuda.synchronize()
    for r in range(P):
        c, sd = ctxs[r], sds[r]
        cb.solver_gaussel_z(c.n, ng, c.hi(), sd.a, bbd, sd.c, cbc[2], cf, alphai, slabs_d[r], arrplan=sd.arrplan, stream=streams[r])
    torch.cuda.synchronize()
    for r, c in enumerate(ctxs):
        assert c.dist_status() == 0
        z0, z1 = SlabDecomp(ng, P, r).z_range()
        gz = slabs_d[r].cpu().numpy()
        assert cases.rel_l2(gz[1:-1, 1:-1, 1:-1], refz[1 + z0:1 + z1, 1:-1, 1:-1]) < tol
    for sd in sds:
        sd.arrplan.destroy()
    for c in ctxs:
        c.close()
    return err


@pytest.mark.parametrize("name", sorted(cases.DIST_CASES))
@pytest.mark.parametrize("P", [2, 3])
def test_virtual_ranks_pipelined(cb, name, P):
    """default: the exchange pipelined over x windows with per-window flags"""
    _run_virtual_ranks(cb, name, P, -1)


@pytest.mark.parametrize("name", ["chan_64x64x64", "uneven_64x64x70", "helm_w_64x64x64"])
def test_virtual_ranks_one_window(cb, name):
    """one window: two whole-field barriers (round 1's schedule)"""
    _run_virtual_ranks(cb, name, 2, 1)


def test_virtual_ranks_four(cb):
    _run_virtual_ranks(cb, "chan_64x64x64", 4, -1)


@pytest.mark.parametrize("name", sorted(cases.DIST_CASES))
@pytest.mark.parametrize("P", [2, 3])
def test_virtual_ranks_copy_engines(cb, name, P):
    """CANSB200_CTX_DIST_MODE = 1: dense local send blocks moved by the copy engines (z chunks forward, x windows back)"""
    _run_virtual_ranks(cb, name, P, -1, mode=1)


@pytest.mark.parametrize("name", sorted(cases.DIST_CASES))
def test_virtual_ranks_two_halves(cb, name):
    """CANSB200_CTX_DIST_MODE = 2: peer stores, z chunks forward (y transforms padded to one CTA per SM), x windows back"""
    _run_virtual_ranks(cb, name, 3 if "fp32" not in name else 2, -1, mode=2)


@pytest.mark.parametrize("windows,chunks", [(1, 1), (2, 3), (8, 8)])
def test_virtual_ranks_copy_engines_shapes(cb, windows, chunks):
    _run_virtual_ranks(cb, "duct_128x64x96", 4, windows, mode=1, chunks=chunks)
    _run_virtual_ranks(cb, "uneven_64x64x70", 3, windows, mode=1, chunks=chunks)
    _run_virtual_ranks(cb, "uneven_64x64x70", 2, windows, mode=2, chunks=chunks)


@pytest.mark.parametrize("tiled", [True, False])
@pytest.mark.parametrize("name,P", [("helm_w_64x64x64", 2), ("uneven_64x64x70", 2), ("helm_w_64x64x64", 3), ("reg_32x64x512", 2),
                                    ("reg_32x64x512", 3), ("m5_32x64x300", 2), ("m3_32x64x160", 2)])
def test_virtual_ranks_distributed_tdma(cb, name, P, tiled):
    """CANSB200_CTX_DTDMA (the reference's is_poisson_dtdma): z stays decomposed, only the reduced system travels; the
    coefficient arrays are cached per coefficient set (the reference's is_dtdma_update / aa_z_save, src/solver_gpu.f90:571-591):
    three solves with two different alpha must run the coefficient kernel twice, not three times."""
    from cans_b200.decomp import SlabDecomp
    ng, l, cbc, cf, gr, dt, helm = cases.DIST_CASES[name]
    dev = torch.device("cuda:0")
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
    p = cases.make_rhs(cs)
    q3 = 1 if (cf[2] == "f" and cbc[2][1] == "D") else 0
    zs = SlabDecomp(ng, P, 0).zs
    ctxs = [cb.Context(ng, rank=r, nranks=P) for r in range(P)]
    cb.Context.connect_local(ctxs)
    for c in ctxs:
        c.set_dtdma(True)
        c.set_dtdma_tiled(tiled)   # slab-local elimination on chip (pipelined kernel) / per-column sweeps in the reference's order
    sds = [cb.initsolver(c, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev, cache_slots=2) for c in ctxs]
    streams = [torch.cuda.Stream() for _ in range(P)]
    slabs_h, slabs_d = [], []
    for r in range(P):
        z0, z1 = zs[r], zs[r + 1]
        h = np.zeros((z1 - z0 + 2, ng[1] + 2, ng[0] + 2), dtype=dt)
        h[1:-1] = p[1 + z0:1 + z1]
        slabs_h.append(h)
        slabs_d.append(torch.from_numpy(h).to(dev))
    for alpha in (cases.ALPHA, 2.5 * cases.ALPHA, cases.ALPHA):
        alphai = dt(1.0) / dt(alpha)
        # oracle: the same stages with the distributed elimination on the global field
        px = np.ascontiguousarray(p[1:-1, 1:-1, 1:-1])
        O.fft(cs["arrplan"][0][0], px)
        O.fft(cs["arrplan"][1][0], px)
        O.gaussel_dtdma(zs, ng[2] - q3, cs["a"], (cs["b"] + alphai).astype(dt), cs["c"], cbc[2] == cases.P, dt(cs["normfft"]) * alphai, px,
                        cs["lambdaxy"])
        O.fft(cs["arrplan"][1][1], px)
        O.fft(cs["arrplan"][0][1], px)
        for r in range(P):
            slabs_d[r].copy_(torch.from_numpy(slabs_h[r]))
        torch.cuda.synchronize()
        for r in range(P):
            c, sd = ctxs[r], sds[r]
            cb.solve_helmholtz(c.n, ng, c.hi(), sd.arrplan, sd.normfft, alpha, sd.lambdaxy, sd.a, sd.b, sd.c, None, None, None,
                               c.is_bound(), cbc, cf, slabs_d[r], stream=streams[r])
        torch.cuda.synchronize()
        for r, c in enumerate(ctxs):
            assert c.dist_status() == 0
            got = slabs_d[r].cpu().numpy()[1:-1, 1:-1, 1:-1]
            assert cases.rel_l2(got, px[zs[r]:zs[r + 1]]) < 1e-12
    for sd in sds:
        assert sd.arrplan.stats()["factorisations"] == 2, sd.arrplan.stats()


def test_dtdma_context_guards(cb):
    """on a CANSB200_CTX_DTDMA context the pivot-cache entry points have nothing to work with (solve_z, gaussel
    -> CANSB200_EUNSUPPORTED instead of a null-pointer write), and the switch is refused once a plan exists."""
    ng, l, cbc, cf, gr, dt, helm = cases.DIST_CASES["helm_w_64x64x64"]
    dev = torch.device("cuda:0")
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
    ctxs = [cb.Context(ng, rank=r, nranks=2) for r in range(2)]
    cb.Context.connect_local(ctxs)
    ctxs[0].set_dtdma(True)
    sd = cb.initsolver(ctxs[0], ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev)
    pd = torch.zeros((ctxs[0].n[2] + 2, ng[1] + 2, ng[0] + 2), dtype=torch.float64, device=dev)
    with pytest.raises(Exception, match="(?i)dtdma"):
        cb.solver_gaussel_z(ctxs[0].n, ng, ctxs[0].hi(), sd.a, sd.b, sd.c, cbc[2], cf, 1.0, pd, arrplan=sd.arrplan)
    import sys
    S = sys.modules["cans_b200.solver"]   # (`cans_b200.solver` the attribute is the function of that name)
    pz = torch.zeros((ctxs[0].n_z[2], ng[1], ng[0]), dtype=torch.float64, device=dev)
    with pytest.raises(Exception, match="(?i)dtdma"):
        S.gaussel(sd.arrplan, ctxs[0].n_z[2], sd.a, sd.b, sd.c, False, 1.0, pz, sd.lambdaxy)
    with pytest.raises(Exception, match="before any plan"):
        ctxs[0].set_dtdma(False)
    sd1 = cb.initsolver(ctxs[1], ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev)
    with pytest.raises(Exception, match="before any plan"):
        ctxs[1].set_dtdma(True)
    del sd1


def test_missing_rank_is_reported(cb):
    """A rank that never calls the collective solve: the device-side wait gives up (no GPU hang), the solve of the
    rank that did call returns normally in device mode (stream ordered, nothing is read back), and the NEXT call on
    that context fails with CANSB200_ECOMM instead of computing on partly written buffers.
    Costs the 20 s device-side time-out once."""
    ng, l, cbc, cf, gr, dt, helm = cases.DIST_CASES["chan_64x64x64"]
    dev = torch.device("cuda:0")
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
    ctxs = [cb.Context(ng, rank=r, nranks=2) for r in range(2)]
    cb.Context.connect_local(ctxs)
    sd = cb.initsolver(ctxs[0], ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev)
    pd = torch.zeros((ctxs[0].n[2] + 2, ng[1] + 2, ng[0] + 2), dtype=torch.float64, device=dev)
    cb.solver(ctxs[0].n, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, pd)   # rank 1 never shows up
    torch.cuda.synchronize()
    assert ctxs[0].dist_status() != 0
    with pytest.raises(Exception, match="timed out"):
        cb.solver(ctxs[0].n, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, pd)
