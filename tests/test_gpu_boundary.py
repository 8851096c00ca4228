"""The drop-in boundary as an UNCHANGED reference host would use it, one test per binding detail:

  * integer plan names for the CUDA build's `integer :: arrplan(2,2)` (src/main.f90:94,107, src/fft.f90:31-35);
  * `solver_gaussel_z(n,ng,hi,a,b,c,bcz,c_or_f,norm,p)` without a plan argument (src/solve_helmholtz.f90:73);
  * eigenvalues in the order an _OPENACC-built `initsolver` produces (src/initsolver.f90:98-117);
  * the three workspace aliases of src/rk.f90:27-29;
  * `updt_rhs_b` with `is_bound` (src/bound.f90:514-598) on the device, on one rank and on z slabs.
"""
import importlib

import numpy as np
import pytest

import cases
from oracle import cans_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
P, N, D = cases.P, cases.N, cases.D


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    return importlib.import_module("cans_b200")


def _dev():
    return torch.device("cuda:0")


def test_plan_ids(cb):
    ctx = cb.Context([16, 16, 16])
    plans = [cb.Plan(ctx, [P, P, N], ["c"] * 3) for _ in range(4)]
    ids = [pl.id for pl in plans]
    assert all(0 < i < 2 ** 31 for i in ids) and len(set(ids)) == 4
    assert [pl.id for pl in plans] == ids                       # stable
    for pl, i in zip(plans, ids):
        assert cb.Plan.handle_from_id(i) == pl.handle.value     # the id names the very same plan
    plans[1].destroy()
    assert cb.Plan.handle_from_id(ids[1]) is None               # a dead plan has no name
    assert cb.Plan.handle_from_id(0) is None and cb.Plan.handle_from_id(10 ** 6) is None
    again = cb.Plan(ctx, [P, P, N], ["c"] * 3)
    assert again.id == ids[1]                                   # ids are recycled, never shared by two live plans
    # an id is enough to solve: what the Fortran shim does with arrplan(1,1)
    cs = O.make_case([16, 16, 16], [1.0, 1.0, 1.0], [P, P, N])
    sd = cb.initsolver(ctx, [16, 16, 16], cs["dli"], cs["dzci"], cs["dzfi"], [P, P, N], cs["bc"], ["c"] * 3, device=_dev())
    assert cb.Plan.handle_from_id(sd.arrplan.id) == sd.arrplan.handle.value


def test_get_work_aliases(cb):
    ctx = cb.Context([32, 16, 8])
    ptrs = []
    for which in (0, 1, 2):
        ptr, nel = ctx.get_work(which)
        assert ptr and nel >= 32 * 16 * 8
        ptrs.append(ptr)
    assert len(set(ptrs)) == 3
    with pytest.raises(Exception):
        ctx.get_work(3)


@pytest.mark.parametrize("name", ["C2s_triperiodic", "C3s_channel", "C4s_duct", "odd_sizes", "prime_17x34", "periodic_z_odd"])
def test_openacc_eigenvalue_order(cb, name):
    """An OpenACC host's unchanged initsolver hands over lambdaxy in the packed order: the plan option lambda_order = 1
    must give the same pressure as the CPU-build order."""
    cs = cases.build_case(name)
    ng, cbc, cf = cs["ng"], cs["cbc"], cs["c_or_f"]
    p = cases.make_rhs(cs)
    ref = cases.oracle_solve(name, cs, p)
    ctx = cb.Context(ng)
    sd = cb.initsolver(ctx, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=_dev(), openacc_order=True)
    lam_pk = sd.lambdaxy.cpu().numpy()
    if any(b == P for b in cbc[:2]) and min(ng[0] if cbc[0] == P else 99, ng[1] if cbc[1] == P else 99) > 2:
        assert not np.array_equal(lam_pk, cs["lambdaxy"]), "the packed order must differ from halfcomplex order"
    assert np.array_equal(np.sort(lam_pk, axis=None), np.sort(cs["lambdaxy"], axis=None))
    pd = torch.from_numpy(p.copy()).to(_dev())
    cb.solver(ng, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, pd)
    torch.cuda.synchronize()
    err = cases.parity_error(cs, pd.cpu().numpy()[1:-1, 1:-1, 1:-1], ref[1:-1, 1:-1, 1:-1])
    assert err < 1e-12, err


@pytest.mark.parametrize("name,nranks", [("tgv_64x128x64", 2), ("chan_64x64x64", 3), ("duct_128x64x96", 2)])
def test_openacc_eigenvalue_order_on_slabs(cb, name, nranks):
    """Several ranks: the rows of a periodic y direction are dealt out to the z pencils in packed order, so that each
    rank's slice lambdaxy(lo_z(1):hi_z(1), lo_z(2):hi_z(2)) of the packed array is the right one (virtual ranks on one GPU)."""
    from cans_b200.decomp import SlabDecomp
    ng, l, cbc, cf, gr, dt, helm = cases.DIST_CASES[name]
    dev = _dev()
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
    p = cases.make_rhs(cs)
    ref = p.copy()
    O.solver(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cbc, cf, ref)
    ctxs = [cb.Context(ng, rank=r, nranks=nranks) for r in range(nranks)]
    cb.Context.connect_local(ctxs)
    sds = [cb.initsolver(c, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev, openacc_order=True) for c in ctxs]
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    slabs = []
    for r in range(nranks):
        z0, z1 = SlabDecomp(ng, nranks, r).z_range()
        h = np.zeros((z1 - z0 + 2, ng[1] + 2, ng[0] + 2), dtype=dt)
        h[1:-1] = p[1 + z0:1 + z1]
        slabs.append(torch.from_numpy(h).to(dev))
    torch.cuda.synchronize()
    for r in range(nranks):
        c, sd = ctxs[r], sds[r]
        cb.solver(c.n, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, slabs[r], stream=streams[r])
    torch.cuda.synchronize()
    assert all(c.dist_status() == 0 for c in ctxs)
    full = np.concatenate([t.cpu().numpy()[1:-1, 1:-1, 1:-1] for t in slabs], axis=0)
    err = cases.parity_error(cs, full, ref[1:-1, 1:-1, 1:-1])
    assert err < 1e-12, err


@pytest.mark.parametrize("bcz,cfz", [(N, "c"), (D, "c"), (D, "f"), (P, "c")])
def test_solver_gaussel_z_reference_signature(cb, bcz, cfz):
    """`call solver_gaussel_z(n,ng,hi,a,bb,c,cbc(:,3),c_or_f,alphai,p)`: ten arguments, no plan."""
    ng = [24, 10, 40]
    cbc, cf = [P, P, bcz], ["c", "c", cfz]
    cs = O.make_case(ng, [1.0, 1.0, 1.0], cbc, c_or_f=cf, gr=0.0 if bcz == P else 1.0)
    p = cases.make_rhs(cs)
    alphai = 1.0 / cases.ALPHA
    bb = cs["b"] + alphai
    ref = p.copy()
    O.solver_gaussel_z(ng, ng, ng, cs["a"], bb, cs["c"], bcz, cf, alphai, ref)
    ctx = cb.Context(ng)
    dev = _dev()
    a, b, c = (torch.from_numpy(v).to(dev) for v in (cs["a"], bb, cs["c"]))
    pd = torch.from_numpy(p.copy()).to(dev)
    cb.solver_gaussel_z(ng, ng, ng, a, b, c, bcz, cf, alphai, pd)          # ten arguments, as in the reference
    cb.solver_gaussel_z(ng, ng, ng, a, b, c, bcz, cf, alphai, pd.clone())  # second call reuses the context's z-only plan
    torch.cuda.synchronize()
    got = pd.cpu().numpy()
    assert cases.rel_l2(got[1:-1, 1:-1, 1:-1], ref[1:-1, 1:-1, 1:-1]) < 1e-12
    halo = np.ones(got.shape, bool)
    halo[1:-1, 1:-1, 1:-1] = False
    assert np.array_equal(got[halo], p[halo])
    # host memory too
    ph = p.copy()
    cb.solver_gaussel_z(ng, ng, ng, cs["a"], bb, cs["c"], bcz, cf, alphai, ph, ctx=ctx)
    assert cases.rel_l2(ph[1:-1, 1:-1, 1:-1], ref[1:-1, 1:-1, 1:-1]) < 1e-12


@pytest.mark.parametrize("cf,cbc", [(["c", "c", "c"], [D, D, D]), (["f", "c", "c"], [D, N, D]), (["c", "c", "f"], [P, P, D]),
                                    (["c", "f", "c"], [N, D, N])])
def test_updt_rhs_b_device(cb, cf, cbc):
    """cansb200_updt_rhs_b against the restatement of src/bound.f90:514-598 (bit-identical: one add per point and wall)."""
    ng = [12, 10, 14]
    rng = np.random.default_rng(5)
    p = rng.uniform(-1, 1, (ng[2] + 2, ng[1] + 2, ng[0] + 2))
    rx, ry, rz = [0.3, -1.7], [2.5, 0.125], [-0.75, 1.1]
    for alpha in (None, cases.ALPHA):
        ref = p.copy()
        O.updt_rhs_b(cf, cbc, ng, rx, ry, rz, ref, alpha)
        ctx = cb.Context(ng)
        pd = torch.from_numpy(p.copy()).to(_dev())
        cb.updt_rhs_b(cf, cbc, ng, ctx.is_bound(), rx, ry, rz, pd, alpha, ctx=ctx)
        torch.cuda.synchronize()
        assert np.array_equal(pd.cpu().numpy(), ref)
        # an absent direction (optional argument not present) is skipped
        ref2 = p.copy()
        O.updt_rhs_b(cf, cbc, ng, None, ry, None, ref2, alpha)
        pd2 = torch.from_numpy(p.copy()).to(_dev())
        cb.updt_rhs_b(cf, cbc, ng, ctx.is_bound(), None, ry, None, pd2, alpha, ctx=ctx)
        torch.cuda.synchronize()
        assert np.array_equal(pd2.cpu().numpy(), ref2)


@pytest.mark.parametrize("nranks", [2, 3])
def test_helmholtz_wall_values_on_slabs(cb, nranks):
    """solve_helmholtz with non-zero wall values on z slabs: the z contributions of updt_rhs_b go to the ranks that own
    the walls only (`is_bound`), and the face-centred Dirichlet top ends one plane early on the LAST rank only."""
    from cans_b200.decomp import SlabDecomp
    ng, l = [32, 64, 36], [1.0, 1.0, 1.0]
    cbc, cf = [P, D, D], ["c", "c", "f"]
    bcv = [[0.0, 0.0], [0.4, -0.3], [1.5, -2.0]]
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=1.0, bc=bcv)
    p = cases.make_rhs(cs)
    ref = p.copy()
    O.solve_helmholtz(ng, ng, cs["arrplan"], cs["normfft"], cases.ALPHA, cs["lambdaxy"], cs["a"], cs["b"], cs["c"],
                      cs["rhsbx"], cs["rhsby"], cs["rhsbz"], cbc, cf, ref)
    dev = _dev()
    ctxs = [cb.Context(ng, rank=r, nranks=nranks) for r in range(nranks)]
    cb.Context.connect_local(ctxs)
    sds = [cb.initsolver(c, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, bcv, cf, device=dev) for c in ctxs]
    assert ctxs[0].is_bound()[2] == [True, False] and ctxs[-1].is_bound()[2] == [False, True]
    streams = [torch.cuda.Stream() for _ in range(nranks)]
    slabs = []
    for r in range(nranks):
        z0, z1 = SlabDecomp(ng, nranks, r).z_range()
        h = np.zeros((z1 - z0 + 2, ng[1] + 2, ng[0] + 2))
        h[1:-1] = p[1 + z0:1 + z1]
        slabs.append(torch.from_numpy(h).to(dev))
    torch.cuda.synchronize()
    for r in range(nranks):
        c, sd = ctxs[r], sds[r]
        np.testing.assert_allclose(sd.rhsbz, cs["rhsbz"])
        cb.solve_helmholtz(c.n, ng, c.hi(), sd.arrplan, sd.normfft, cases.ALPHA, sd.lambdaxy, sd.a, sd.b, sd.c, sd.rhsbx, sd.rhsby,
                           sd.rhsbz, c.is_bound(), cbc, cf, slabs[r], stream=streams[r])
    torch.cuda.synchronize()
    assert all(c.dist_status() == 0 for c in ctxs)
    full = np.concatenate([t.cpu().numpy()[1:-1, 1:-1, 1:-1] for t in slabs], axis=0)
    assert cases.rel_l2(full, ref[1:-1, 1:-1, 1:-1]) < 1e-12


@pytest.mark.parametrize("name", ["C3s_channel", "C4s_duct", "helm_w_face_z", "periodic_z_odd"])
def test_reference_call_sequence(cb, name):
    """fftini(ng,n_x,n_y,bcxy,c_or_f,arrplan,normfft) -> solver(n,ng,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p) ->
    fftend(arrplan) with an INTEGER arrplan(1,1), exactly the three calls (and the argument lists) of the reference:
    cansb200_fftini / cansb200_solver / cansb200_fftend through ctypes, as the Fortran shim binds them."""
    import ctypes as C
    from cans_b200._lib import lib, check, i3
    cs = cases.build_case(name)
    ng, cbc, cf = cs["ng"], cs["cbc"], cs["c_or_f"]
    helm = name in cases.HELMHOLTZ
    p = cases.make_rhs(cs)
    ref = cases.oracle_solve(name, cs, p, helm)
    ctx = cb.Context(ng)
    bcxy = "".join(cbc[d][i] for d in range(2) for i in range(2)).encode()
    nf, pid = C.c_double(), C.c_int()
    check(lib.cansb200_fftini(ctx.handle, bcxy, "".join(cf[:2]).encode(), C.byref(nf), C.byref(pid)), "fftini")
    assert pid.value >= 1 and nf.value == pytest.approx(float(cs["normfft"]), rel=1e-15)
    dev = _dev()
    lam, a, b, c = (torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (cs["lambdaxy"], cs["a"], cs["b"], cs["c"]))
    normfft = nf.value
    if helm:   # solve_helmholtz's own arithmetic (src/solve_helmholtz.f90:63-71) stays on the host side
        alphai = 1.0 / cases.ALPHA
        b = b + alphai
        normfft = normfft * alphai
    bc6 = "".join(cbc[d][i] for d in range(3) for i in range(2)).encode()
    pd = torch.from_numpy(p.copy()).to(dev)
    for _ in range(2):
        pd.copy_(torch.from_numpy(p))
        check(lib.cansb200_solver(ctx.handle, pid.value, bc6, "".join(cf).encode(), pd.data_ptr(), i3(ng), 1, normfft, lam.data_ptr(),
                                  a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, 1, torch.cuda.current_stream().cuda_stream), "solver")
    torch.cuda.synchronize()
    err = cases.parity_error(cs, pd.cpu().numpy()[1:-1, 1:-1, 1:-1], ref[1:-1, 1:-1, 1:-1], helm)
    assert err < 1e-12, err
    # x / y boundary conditions that differ from fftini's are refused, as is a dead id
    bad = (b"NN" + bc6[2:]) if bc6[:2] != b"NN" else (b"DD" + bc6[2:])
    assert lib.cansb200_solver(ctx.handle, pid.value, bad, "".join(cf).encode(), pd.data_ptr(), i3(ng), 1, normfft, lam.data_ptr(),
                               a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, 1, None) != 0
    check(lib.cansb200_fftend(ctx.handle, pid.value), "fftend")
    assert lib.cansb200_fftend(ctx.handle, pid.value) != 0
    assert lib.cansb200_solver(ctx.handle, pid.value, bc6, "".join(cf).encode(), pd.data_ptr(), i3(ng), 1, normfft, lam.data_ptr(),
                               a.data_ptr(), b.data_ptr(), c.data_ptr(), 0, 1, None) != 0
