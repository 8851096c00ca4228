"""GPU: the pressure-correction right-hand side fused into the solve (`cansb200_solve_fillps`).

Reference sequence being replaced by one call (src/main.f90:465-467):
    call fillps(n,dli,dzfi,dtrki,u,v,w,pp)                                   src/fillps.f90:13-51
    call updt_rhs_b(['c','c','c'],cbcpre,n,is_bound,rhsbp%x,rhsbp%y,rhsbp%z,pp)   src/bound.f90:514-598
    call solver(n,ng,arrplanp,normfftp,lambdaxyp,ap,bp,cp,cbcpre,['c','c','c'],pp)
The forward x transform evaluates fillps (+ the wall terms) at load time from u, v, w.  Checked against the oracle's
fillps -> updt_rhs_b -> solver on identical inputs (1e-12 FP64 / 1e-5 FP32) and against the unfused CUDA sequence."""
import importlib

import numpy as np
import pytest

import cases
from oracle import cans_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

P, N, D = cases.P, cases.N, cases.D
C3 = ["c", "c", "c"]


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    return importlib.import_module("cans_b200")


def _velocity(ng, dtype, seed=7):
    """u, v, w on the haloed grid, halos included (fillps reads u(i-1), v(j-1), w(k-1) of the first interior cells)"""
    hg = [ng[0] + 2, ng[1] + 2, ng[2] + 2]
    return [(0.5 * O.hash_field(hg, seed + s)).astype(dtype) for s in range(3)]


def _oracle(cs, uvw, dti, rh):
    ng, dt = cs["ng"], cs["dtype"]
    ref = np.zeros((ng[2] + 2, ng[1] + 2, ng[0] + 2), dtype=dt)
    O.fillps(ng, cs["dli"], cs["dzfi"], dti, uvw[0], uvw[1], uvw[2], ref)
    if rh is not None:
        O.updt_rhs_b(cs["c_or_f"], cs["cbc"], ng, rh[0], rh[1], rh[2], ref)
    O.solver(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cs["cbc"], cs["c_or_f"], ref)
    return ref


# name -> (ng, cbc, dtype, wall values or None, expect the fused kernel)
FUSED_CASES = {
    "r2hc_64_odd_lines": ([64, 5, 7], [P, P, D], np.float64, None, 1),          # 35 lines: the last pair has no second line
    "redft10_128": ([128, 6, 9], [N, P, D], np.float64, None, 1),
    "rodft10_256": ([256, 4, 6], [D, N, D], np.float64, None, 1),
    "r2hc_384": ([384, 3, 4], [P, N, D], np.float64, None, 1),
    "r2hc_512": ([512, 4, 5], [P, P, D], np.float64, None, 1),
    "redft10_768": ([768, 2, 2], [N, N, D], np.float64, None, 1),
    "r2hc_1024": ([1024, 2, 3], [P, P, D], np.float64, None, 1),
    "r2hc_2048": ([2048, 2, 2], [P, P, D], np.float64, None, 1),
    "walls_rodft10_64": ([64, 6, 10], [D, N, D], np.float64, ([0.3, -0.7], [0.11, 0.05], [-0.4, 0.9]), 1),
    "walls_r2hc_128": ([128, 4, 6], [P, N, D], np.float64, (None, [0.2, -0.1], [0.6, 0.25]), 1),
    "fp32_r2hc_128": ([128, 8, 8], [P, P, D], np.float32, None, 1),
    "fp32_walls_redft10_64": ([64, 6, 6], [N, D, D], np.float32, ([0.5, 0.25], [0.3, -0.2], None), 1),
    "generic_x_48": ([48, 6, 8], [P, P, D], np.float64, ([0.0, 0.0], None, [0.4, -0.3]), 0),   # no two-for-one kernel: three steps
    "singular_channel_64": ([64, 16, 24], [P, P, N], np.float64, None, 1),     # the Poisson operator of C3 (pinned null mode)
}


@pytest.mark.parametrize("name", sorted(FUSED_CASES))
def test_solver_fillps_matches_oracle(cb, name):
    ng, cbc, dt, rh, expect = FUSED_CASES[name]
    l = [2.0, 1.0, 1.5]
    cs = O.make_case(ng, l, cbc, gr=1.0 if name != "singular_channel_64" else 0.0, dtype=dt)
    dev = torch.device("cuda:0")
    uvw = _velocity(ng, dt)
    if name == "singular_channel_64":
        # a compatible right-hand side: periodic halos in x, y and no flow through the z walls
        for t in uvw:
            t[:, :, 0] = t[:, :, -2]; t[:, :, -1] = t[:, :, 1]
            t[:, 0, :] = t[:, -2, :]; t[:, -1, :] = t[:, 1, :]
        uvw[2][0] = 0.0; uvw[2][-2] = 0.0
    dti = 1.0 / 0.37
    ref = _oracle(cs, uvw, dti, rh)
    ctx = cb.Context(ng, is_fp32=dt == np.float32)
    sd = cb.initsolver(ctx, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], C3, device=dev)
    u, v, w = (torch.from_numpy(t).to(dev) for t in uvw)
    dzfi = torch.from_numpy(cs["dzfi"]).to(dev)
    shp = (ng[2] + 2, ng[1] + 2, ng[0] + 2)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    p = torch.full(shp, 3.25, dtype=tdt, device=dev)     # the input content of p must not matter
    kw = {} if rh is None else dict(is_bound=ctx.is_bound(), rhsbx=rh[0], rhsby=rh[1], rhsbz=rh[2])
    S = importlib.import_module("cans_b200.solver")
    for rep in range(2):   # second call: pivot-cache hit, tables cached
        S.solver_fillps(ng, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, C3, cs["dli"], dzfi, dti, u, v, w, p, **kw)
    torch.cuda.synchronize()
    assert sd.arrplan.stats()["fillps_fused"] == expect
    got = p.cpu().numpy()
    tol = 1e-12 if dt == np.float64 else 1e-5
    err = cases.parity_error(cs, got[1:-1, 1:-1, 1:-1], ref[1:-1, 1:-1, 1:-1])
    assert err < tol, f"{name}: fused vs oracle rel L2 {err:.3e}"
    halo = np.ones(shp, bool)
    halo[1:-1, 1:-1, 1:-1] = False
    assert np.all(got[halo] == 3.25), "halo cells of p were modified"
    for t, h in zip((u, v, w), uvw):
        assert np.array_equal(t.cpu().numpy(), h), "u, v, w are inputs"
    # the unfused CUDA sequence on the same context: fillps, updt_rhs_b, solver
    ctx.set_fuse_fillps(False)
    p2 = torch.zeros(shp, dtype=tdt, device=dev)
    S.solver_fillps(ng, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, C3, cs["dli"], dzfi, dti, u, v, w, p2, **kw)
    torch.cuda.synchronize()
    assert sd.arrplan.stats()["fillps_fused"] == 0
    err2 = cases.parity_error(cs, got[1:-1, 1:-1, 1:-1], p2.cpu().numpy()[1:-1, 1:-1, 1:-1])
    assert err2 < (1e-13 if dt == np.float64 else 1e-5), f"{name}: fused vs unfused CUDA rel L2 {err2:.3e}"
    p3 = torch.zeros(shp, dtype=tdt, device=dev)
    S.fillps(ctx, ng, cs["dli"], dzfi, dti, u, v, w, p3)
    if rh is not None:
        S.updt_rhs_b(C3, cbc, ng, ctx.is_bound(), rh[0], rh[1], rh[2], p3, ctx=ctx)
    cb.solver(ng, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, C3, p3)
    torch.cuda.synchronize()
    assert torch.equal(p2, p3), "the fallback of solver_fillps is exactly fillps + updt_rhs_b + solver"
    sd.arrplan.destroy()
    ctx.close()


@pytest.mark.parametrize("nranks,walls", [(2, False), (3, True)])
def test_solver_fillps_on_z_slabs(cb, nranks, walls):
    """The fused source under the z-slab decomposition (virtual ranks on one GPU): every rank evaluates the divergence of
    its own slab (halo planes from its neighbours), wall terms only where `is_bound` says so."""
    from cans_b200.decomp import SlabDecomp
    S = importlib.import_module("cans_b200.solver")
    ng, l, cbc, dt = [64, 64, 48], [6.0, 3.0, 2.0], [P, N, D], np.float64
    cs = O.make_case(ng, l, cbc, gr=1.5, dtype=dt)
    dev = torch.device("cuda:0")
    uvw = _velocity(ng, dt, seed=31)
    rh = ([0.0, 0.0], [0.3, -0.2], [0.7, -0.45]) if walls else None
    dti = 2.5
    ref = _oracle(cs, uvw, dti, rh)
    ctxs = [cb.Context(ng, rank=r, nranks=nranks) for r in range(nranks)]
    cb.Context.connect_local(ctxs)
    sds = [cb.initsolver(c, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], C3, device=dev) for c in ctxs]
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    slabs, ps, dz = [], [], []
    for r in range(nranks):
        z0, z1 = SlabDecomp(ng, nranks, r).z_range()
        slabs.append([torch.from_numpy(np.ascontiguousarray(t[z0:z1 + 2])).to(dev) for t in uvw])
        dz.append(torch.from_numpy(np.ascontiguousarray(cs["dzfi"][z0:z1 + 2])).to(dev))
        ps.append(torch.zeros((z1 - z0 + 2, ng[1] + 2, ng[0] + 2), dtype=torch.float64, device=dev))
    torch.cuda.synchronize()
    for rep in range(2):
        for r in range(nranks):   # every rank's solve is enqueued before anything synchronises
            c, sd = ctxs[r], sds[r]
            kw = {} if rh is None else dict(is_bound=c.is_bound(), rhsbx=rh[0], rhsby=rh[1], rhsbz=rh[2])
            S.solver_fillps(c.n, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, C3, cs["dli"], dz[r], dti,
                            slabs[r][0], slabs[r][1], slabs[r][2], ps[r], stream=streams[r], **kw)
        torch.cuda.synchronize()
    for c, sd in zip(ctxs, sds):
        assert c.dist_status() == 0, "a device-side wait timed out"
        assert sd.arrplan.stats()["fillps_fused"] == 1
    full = np.concatenate([t.cpu().numpy()[1:-1, 1:-1, 1:-1] for t in ps], axis=0)
    err = cases.rel_l2(full, ref[1:-1, 1:-1, 1:-1])
    assert err < 1e-12, f"P={nranks}: fused vs oracle rel L2 {err:.3e}"
    for sd in sds:
        sd.arrplan.destroy()
    for c in ctxs:
        c.close()


def test_solver_fillps_argument_errors(cb):
    S = importlib.import_module("cans_b200.solver")
    ng, cbc = [64, 4, 4], [P, P, D]
    cs = O.make_case(ng, [1.0, 1.0, 1.0], cbc)
    dev = torch.device("cuda:0")
    ctx = cb.Context(ng)
    sd = cb.initsolver(ctx, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], C3, device=dev)
    shp = (ng[2] + 2, ng[1] + 2, ng[0] + 2)
    u, v, w, p = (torch.zeros(shp, dtype=torch.float64, device=dev) for _ in range(4))
    dzfi = torch.from_numpy(cs["dzfi"]).to(dev)
    args = (ng, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, C3, cs["dli"], dzfi, 1.0)
    with pytest.raises(ValueError):
        S.solver_fillps(*args, u.cpu(), v, w, p)                      # host array
    with pytest.raises(ValueError):
        S.solver_fillps(*args, u[:-1], v, w, p)                       # wrong shape
    with pytest.raises(ValueError):
        S.solver_fillps(*args, u, v, w, p, rhsbz=[1.0, 2.0])          # wall values without is_bound
    from cans_b200._lib import lib, i3, d3
    rc = lib.cansb200_solve_fillps(sd.arrplan.handle, p.data_ptr(), i3(ng), 1, 1.0, sd.lambdaxy.data_ptr(), sd.a.data_ptr(),
                                   sd.b.data_ptr(), sd.c.data_ptr(), d3(cs["dli"]), dzfi.data_ptr(), 1.0, None, v.data_ptr(),
                                   w.data_ptr(), None, None, None, None)
    assert rc == -1   # CANSB200_EINVAL: null u
    sd.arrplan.destroy()
    ctx.close()
