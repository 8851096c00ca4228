"""CPU tests that pin the oracle: the reference's own in-binary self-test
properties (src/sanity.f90:206-409) restated -- divergence after the pressure
correction < `small`, Helmholtz residual < `small` -- plus the committed golden
vectors and the host-side initsolver mirror."""
import os

import numpy as np
import pytest

import cases
from oracle import cans_oracle as O

P, N, D = cases.P, cases.N, cases.D


def _sanity_divergence(ng, l, cbcpre, cbcvel, gr, dtype):
    cs = O.make_case(ng, l, cbcpre, gr=gr, dtype=dtype)
    shp = (ng[2] + 2, ng[1] + 2, ng[0] + 2)
    u, v, w, p = (np.zeros(shp, dtype) for _ in range(4))
    I = (slice(1, -1),) * 3
    u[I] = 0.5 * O.hash_field(ng, 123, dtype)
    v[I] = 0.5 * O.hash_field(ng, 456, dtype)
    w[I] = 0.5 * O.hash_field(ng, 789, dtype)
    dl = [1 / x for x in cs["dli"]]
    bcvel = [[[0.0, 0.0]] * 3] * 3
    O.bounduvw(cbcvel, ng, bcvel, dl, cs["dzc"], cs["dzf"], u, v, w)
    dt = np.arccos(-1.0)
    O.fillps(ng, cs["dli"], cs["dzfi"], 1 / dt, u, v, w, p)
    O.updt_rhs_b(["c"] * 3, cbcpre, ng, cs["rhsbx"], cs["rhsby"], cs["rhsbz"], p)
    O.solver(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cbcpre, ["c"] * 3, p)
    O.boundp(cbcpre, ng, [[0.0, 0.0]] * 3, dl, cs["dzc"], p)
    O.correc(ng, cs["dli"], cs["dzci"], dt, p, u, v, w)
    O.bounduvw(cbcvel, ng, bcvel, dl, cs["dzc"], cs["dzf"], u, v, w, keep_norm_values=True)
    return O.chkdiv(ng, l, cs["dli"], cs["dzfi"], u, v, w)[1]


@pytest.mark.parametrize("ng,l,cbcpre,cbcvel,gr,dtype", [
    ([2, 64, 64], [0.03125, 1, 1], [P, N, N], [[P, D, D]] * 3, 0.0, np.float64),      # lid-driven cavity input.nml
    ([32, 32, 32], [6.28, 6.28, 6.28], [P, P, P], [[P, P, P]] * 3, 0.0, np.float64),  # triperiodic
    ([32, 16, 48], [12, 6, 2], [P, P, N], [[P, P, D]] * 3, 2.0, np.float64),          # channel, stretched
    ([32, 24, 48], [12, 2, 2], [P, N, N], [[P, D, D]] * 3, 1.5, np.float64),          # periodic duct
    ([30, 27, 45], [12, 6, 2], [P, N, N], [[P, D, D]] * 3, 1.5, np.float64),          # non power-of-two
    ([16, 24, 20], [1, 1, 1], [N, ["N", "D"], D], [[D, ["D", "N"], N]] * 3, 1.0, np.float64),  # outflow-like
    ([32, 24, 48], [12, 2, 2], [P, N, N], [[P, D, D]] * 3, 1.5, np.float32),
])
def test_pressure_correction_is_divergence_free(ng, l, cbcpre, cbcvel, gr, dtype):
    """src/sanity.f90:262-283: divmax < small."""
    assert _sanity_divergence(ng, l, cbcpre, cbcvel, gr, dtype) < O.small(dtype)


@pytest.mark.parametrize("name", ["helm_u_face_x", "helm_v_face_y", "helm_w_face_z"])
def test_helmholtz_residual(name):
    """src/sanity.f90:284-401 with chk_helmholtz (src/debug.f90:16-90): resmax < small."""
    cs = cases.build_case(name)
    ng = cs["ng"]
    p = cases.make_rhs(cs)
    rhs = p.copy()
    ref = cases.oracle_solve(name, cs, p, helmholtz=True)
    dl = [1 / x for x in cs["dli"]]
    for idir in range(3):
        for ib in (0, 1):
            centered = cs["c_or_f"][idir] == "c"
            if idir < 2:
                dr = dl[idir]
            else:
                dr = (cs["dzc"] if centered else cs["dzf"])[0 if ib == 0 else ng[2]]
            O.set_bc(cs["cbc"][idir][ib], ib, idir, centered, 0.0, dr, ref)
    res = O.chk_helmholtz(ng, cs["l"], cs["dli"], cs["dzci"], cs["dzfi"], 1 / cases.ALPHA, rhs / cases.ALPHA, ref,
                          cs["cbc"], cs["c_or_f"])
    assert res[1] < O.small(np.float64)


@pytest.mark.parametrize("name", [f[:-4] for f in sorted(os.listdir(os.path.join(os.path.dirname(__file__), "golden")))
                                  if f.endswith(".npz")])
def test_oracle_reproduces_golden(name):
    cs = cases.build_case(name)
    ref = cases.oracle_solve(name, cs, cases.make_rhs(cs), helmholtz=name in cases.HELMHOLTZ)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))["p"]
    tol = 1e-13 if cs["dtype"] == np.float64 else 1e-6
    assert cases.rel_l2(ref[1:-1, 1:-1, 1:-1], gold) < tol


def test_r2r_definitions_match_fftw_manual():
    """The scipy-based r2r restatement against a direct evaluation of the FFTW manual's sums."""
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 8, 9):
        x = rng.standard_normal(n)
        j = np.arange(n)
        k = j[:, None]
        want = {
            O.FFTW_REDFT10: 2 * (x * np.cos(np.pi * (j + 0.5) * k / n)).sum(1),
            O.FFTW_REDFT01: x[0] + 2 * (x[1:] * np.cos(np.pi * j[1:] * (k + 0.5) / n)).sum(1),
            O.FFTW_RODFT10: 2 * (x * np.sin(np.pi * (j + 0.5) * (k + 1) / n)).sum(1),
            O.FFTW_RODFT01: (-1.0) ** j * x[n - 1] + 2 * (x[:-1] * np.sin(np.pi * (j[:-1] + 1) * (k + 0.5) / n)).sum(1),
            O.FFTW_REDFT11: 2 * (x * np.cos(np.pi * (j + 0.5) * (k + 0.5) / n)).sum(1),
            O.FFTW_RODFT11: 2 * (x * np.sin(np.pi * (j + 0.5) * (k + 0.5) / n)).sum(1),
            O.FFTW_RODFT00: 2 * (x * np.sin(np.pi * (j + 1) * (k + 1) / (n + 1))).sum(1),
        }
        if n > 1:
            want[O.FFTW_REDFT00] = x[0] + (-1.0) ** j * x[n - 1] + 2 * (x[1:-1] * np.cos(np.pi * j[1:-1] * k / (n - 1))).sum(1)
        for kind, w in want.items():
            np.testing.assert_allclose(O.r2r_1d(x, kind), w, rtol=0, atol=1e-12)
        X = np.fft.rfft(x)
        hc = O.r2r_1d(x, O.FFTW_R2HC)
        np.testing.assert_allclose(hc[:n // 2 + 1], X.real, atol=1e-13)
        for f in range(1, (n + 1) // 2):
            assert abs(hc[n - f] - X.imag[f]) < 1e-13
        np.testing.assert_allclose(O.r2r_1d(hc, O.FFTW_HC2R), n * x, atol=1e-12)


def test_host_initsolver_matches_oracle():
    """cans_b200.solver.initsolver pieces (host arithmetic) against the oracle's restatement."""
    import importlib
    S = importlib.import_module("cans_b200.solver")
    for bc in (P, N, D, ["N", "D"], ["D", "N"]):
        for cf in "cf":
            for n in (2, 9, 64):
                np.testing.assert_array_equal(S.eigenvalues(n, bc, cf), O.eigenvalues(n, bc, cf))
                assert S.find_fft(bc, cf) == O.find_fft(bc, cf)
            dzc, dzf, _, _ = O.initgrid(1, 24, 1.3, 2.0, bc == P)
            for got, want in zip(S.tridmatrix(bc, 24, 1 / dzc, 1 / dzf, cf), O.tridmatrix(bc, 24, 1 / dzc, 1 / dzf, cf)):
                np.testing.assert_array_equal(got, want)
            assert S.bc_rhs(bc, [0.3, -0.7], [0.1, 0.2], [0.3, 0.4], cf) == O.bc_rhs(bc, [0.3, -0.7], [0.1, 0.2], [0.3, 0.4], cf)


def test_gaussel_pins_singular_mode():
    """src/solver.f90:151-164: the lambda = 0 column of an all-Neumann problem is pinned, not regularised."""
    cs = cases.build_case("C1_ldc_2x64x64")
    p = cases.make_rhs(cs)
    ref = cases.oracle_solve("C1_ldc_2x64x64", cs, p)
    assert np.isfinite(ref).all()
    assert cs["lambdaxy"][0, 0] == 0.0


@pytest.mark.parametrize("name", ["C3s_channel", "periodic_z_odd", "helm_w_face_z", "tiny_z"])
def test_solver_gaussel_z_against_dense_solve(name):
    """The oracle of solver_gaussel_z (src/solver.f90:547-616, lambda-less gaussel) against a dense solve of the
    same tridiagonal (cyclic for periodic z) system (b shifted by 1/alpha as the implicit z diffusion does)."""
    cs = cases.build_case(name)
    ng = cs["ng"]
    a, c = cs["a"].astype(np.float64), cs["c"].astype(np.float64)
    b = cs["b"].astype(np.float64) + 1.0 / cases.ALPHA
    per = cs["cbc"][2] == P
    q = 1 if (cs["c_or_f"][2] == "f" and cs["cbc"][2][1] == "D") else 0
    n = ng[2] - q
    rng = np.random.default_rng(17)
    p = rng.uniform(-1, 1, (ng[2] + 2, ng[1] + 2, ng[0] + 2))
    ref = p.copy()
    norm = 0.37
    O.solver_gaussel_z(ng, ng, ng, a, b, c, cs["cbc"][2], cs["c_or_f"], norm, ref)
    M = np.zeros((n, n))
    for k in range(n):
        M[k, k] = b[k]
        if k > 0:
            M[k, k - 1] = a[k]
        if k < n - 1:
            M[k, k + 1] = c[k]
    if per:
        M[0, n - 1] += a[0]
        M[n - 1, 0] += c[n - 1]
    rhs = p[1:n + 1, 1:-1, 1:-1].reshape(n, -1) * norm
    want = np.linalg.solve(M, rhs).reshape(n, ng[1], ng[0])
    got = ref[1:n + 1, 1:-1, 1:-1]
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-11
    # untouched: halos and, for a face-centred Dirichlet top, the last plane
    mask = np.ones(p.shape, bool)
    mask[1:n + 1, 1:-1, 1:-1] = False
    assert np.array_equal(ref[mask], p[mask])


@pytest.mark.parametrize("name,nranks", [("C3s_channel", 2), ("C3s_channel", 4), ("C4s_duct", 3), ("periodic_z_odd", 3),
                                         ("C2s_triperiodic", 4), ("helm_w_face_z", 2)])
def test_gaussel_dtdma_oracle(name, nranks):
    """The distributed-TDMA restatement (src/solver.f90:309-517, `is_poisson_dtdma`) solves the same tridiagonal
    systems as `gaussel` and as a dense solve, for even and uneven z splits and for periodic z."""
    from cans_b200.decomp import split_starts
    cs = cases.build_case(name)
    ng = cs["ng"]
    helm = name in cases.HELMHOLTZ
    a, c = cs["a"], cs["c"]
    b = cs["b"] + (1.0 / cases.ALPHA if helm else 0.0)
    lam = cs["lambdaxy"] - (0.0 if helm or not cases.is_singular(cs) else 0.37)   # keep every column regular
    per = cs["cbc"][2] == P
    q = 1 if (cs["c_or_f"][2] == "f" and cs["cbc"][2][1] == "D") else 0
    n = ng[2] - q
    rng = np.random.default_rng(23)
    pz = rng.uniform(-1, 1, (ng[2], ng[1], ng[0]))
    norm = 0.61
    ref = pz.copy()
    O.gaussel(n, a, b, c, per, norm, ref, lam)
    got = pz.copy()
    O.gaussel_dtdma(split_starts(ng[2], nranks), n, a, b, c, per, norm, got, lam)
    assert per == (name in ("periodic_z_odd", "C2s_triperiodic"))
    scale = np.abs(ref[:n]).max()
    assert np.abs(got[:n] - ref[:n]).max() / scale < 1e-10
    assert np.array_equal(got[n:], pz[n:])
    # dense check of one column
    j, i = ng[1] // 2, ng[0] // 3
    M = np.zeros((n, n))
    for k in range(n):
        M[k, k] = b[k] + lam[j, i]
        if k > 0:
            M[k, k - 1] = a[k]
        if k < n - 1:
            M[k, k + 1] = c[k]
    if per:
        M[0, n - 1] += a[0]
        M[n - 1, 0] += c[n - 1]
    want = np.linalg.solve(M, pz[:n, j, i] * norm)
    assert np.abs(got[:n, j, i] - want).max() / np.abs(want).max() < 1e-10
