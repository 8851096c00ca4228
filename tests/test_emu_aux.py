"""The step kernels' own source (cans_b200/csrc/aux_kernels.cuh), compiled by g++ and run thread by thread on the CPU
(tests/emu/emu_aux.cpp) with the launch geometry the library uses, against the oracle: fillps (src/fillps.f90:38-50), correc
(src/correc.f90:33-59), updt_rhs_b (src/bound.f90:514-598), the eigenvalue reordering of an _OPENACC-built initsolver
(src/initsolver.f90:98-117) and the synthetic hash field.  Bit-exact in both precisions: same expressions, same order."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cans_oracle as O

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "emu_aux")
SHAPES = [([2, 64, 64], np.float64), ([37, 5, 9], np.float64), ([300, 3, 4], np.float64), ([33, 17, 6], np.float32)]
IDS = [f"{s[0][0]}x{s[0][1]}x{s[0][2]}_{np.dtype(s[1]).name}" for s in SHAPES]


def _emu(dt, op, ng, d, *params):
    if not os.path.exists(EMU):   # built by __graft_entry__.build() (conftest's session fixture)
        pytest.skip("tests/_build/emu_aux was not built (g++ or the CUDA headers are missing)")
    prec = "f32" if dt == np.float32 else "f64"
    r = subprocess.run([EMU, prec, op, str(ng[0]), str(ng[1]), str(ng[2]), str(d)] + [repr(float(p)) if not isinstance(p, (int, np.integer)) else str(int(p)) for p in params],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def _fields(ng, dt, d, names="uvwp"):
    hg = [ng[0] + 2, ng[1] + 2, ng[2] + 2]
    out = {}
    for s, nm in enumerate(names):
        out[nm] = (0.5 * O.hash_field(hg, 70 + s)).astype(dt)
        out[nm].tofile(os.path.join(d, nm + ".bin"))
    return out


def _read(d, name, shape, dt):
    return np.fromfile(os.path.join(d, name + ".bin"), dtype=dt).reshape(shape)


@pytest.mark.parametrize("op", ["fillps3d", "fillps_flat"])
@pytest.mark.parametrize("ng,dt", SHAPES, ids=IDS)
def test_fillps_kernels(tmp_path, ng, dt, op):
    cs = O.make_case(ng, [2.0, 1.0, 1.5], [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0, dtype=dt)
    f = _fields(ng, dt, tmp_path)
    cs["dzfi"].tofile(os.path.join(tmp_path, "dzfi.bin"))
    dti = dt(1.0 / 0.37)
    ref = f["p"].copy()
    O.fillps(ng, cs["dli"], cs["dzfi"], dti, f["u"], f["v"], f["w"], ref)
    _emu(dt, op, ng, tmp_path, cs["dli"][0], cs["dli"][1], dti)
    assert np.array_equal(_read(tmp_path, "p_out", ref.shape, dt), ref)


@pytest.mark.parametrize("op", ["correc3d", "correc_flat"])
@pytest.mark.parametrize("ng,dt", SHAPES, ids=IDS)
def test_correc_kernels(tmp_path, ng, dt, op):
    cs = O.make_case(ng, [2.0, 1.0, 1.5], [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0, dtype=dt)
    f = _fields(ng, dt, tmp_path)
    cs["dzci"].tofile(os.path.join(tmp_path, "dzci.bin"))
    dtc = dt(0.37)
    ur, vr, wr = f["u"].copy(), f["v"].copy(), f["w"].copy()
    O.correc(ng, cs["dli"], cs["dzci"], dtc, f["p"], ur, vr, wr)
    _emu(dt, op, ng, tmp_path, cs["dli"][0], cs["dli"][1], dtc)
    for nm, r in (("u_out", ur), ("v_out", vr), ("w_out", wr)):
        assert np.array_equal(_read(tmp_path, nm, r.shape, dt), r), nm


@pytest.mark.parametrize("cf,cbc", [(["c", "c", "c"], [["D", "D"], ["N", "N"], ["D", "N"]]),
                                    (["f", "c", "f"], [["D", "D"], ["P", "P"], ["N", "D"]])])
@pytest.mark.parametrize("ng,dt", SHAPES[1:], ids=IDS[1:])
def test_updt_rhs_b_kernel(tmp_path, ng, dt, cf, cbc):
    f = _fields(ng, dt, tmp_path, "p")
    rh = [[0.3, -0.7], [0.11, 0.05], [-0.4, 0.9]]
    alpha = -0.0123
    ref = f["p"].copy()
    O.updt_rhs_b(cf, cbc, ng, rh[0], rh[1], rh[2], ref, alpha)
    q = [1 if (cf[d] == "f" and cbc[d][1] == "D") else 0 for d in range(3)]
    idx = [v for d in range(3) for v in (1, ng[d] - q[d])]
    vals = [float(dt(rh[d][s]) * dt(alpha)) for d in range(3) for s in range(2)]   # rhsb * norm in the working precision
    _emu(dt, "updt_rhs_b", ng, tmp_path, *idx, *vals)
    assert np.array_equal(_read(tmp_path, "p_out", ref.shape, dt), ref)


@pytest.mark.parametrize("nx,ny", [(8, 6), (9, 7), (64, 5), (2, 2), (1, 3)])
def test_lambda_unpack_kernel(tmp_path, nx, ny):
    """lambda in the packed order of an _OPENACC-built initsolver -> the order the kernels keep the spectrum in"""
    from cans_b200.solver import eigenvalues
    lx_hc, lx_pk = eigenvalues(nx, "PP", "c"), eigenvalues(nx, "PP", "c", openacc_order=True)
    ly_hc, ly_pk = eigenvalues(ny, "PP", "c") * 3.0, eigenvalues(ny, "PP", "c", openacc_order=True) * 3.0
    want = ly_hc[:, None] + lx_hc[None, :]
    for px, py in ((1, 1), (1, 0), (0, 1), (0, 0)):
        lam = (ly_pk if py else ly_hc)[:, None] + (lx_pk if px else lx_hc)[None, :]
        np.ascontiguousarray(lam).tofile(os.path.join(tmp_path, "lam.bin"))
        _emu(np.float64, "lambda_unpack", [nx, ny, 1], tmp_path, px, py, 0)
        # the two halves of a periodic spectrum share their eigenvalue up to the rounding of cos(2 pi (n - l) / n) against cos(2 pi l / n)
        np.testing.assert_allclose(_read(tmp_path, "lam_out", (ny, nx), np.float64), want, rtol=1e-14, atol=1e-14, err_msg=str((px, py)))
    if nx % 2 == 0 and nx >= 2:
        # split order inside the solve: (r0 .. r[n/2-1] | r[n/2], i1 .. i[n/2-1]); halfcomplex position of split position s
        hc_of = [s if 2 * s <= nx else nx - (s - nx // 2) for s in range(nx)]
        np.ascontiguousarray(want).tofile(os.path.join(tmp_path, "lam.bin"))
        _emu(np.float64, "lambda_unpack", [nx, ny, 1], tmp_path, 0, 0, 1)
        assert np.array_equal(_read(tmp_path, "lam_out", (ny, nx), np.float64), want[:, hc_of])


@pytest.mark.parametrize("nhalo", [0, 1])
def test_fill_hash_kernel(tmp_path, nhalo):
    """the device twin of oracle.hash_field: indexed by the GLOBAL (i, j, k), so any decomposition sees the same field"""
    ngl, n, lo = [12, 7, 9], [12, 7, 4], [0, 0, 3]     # a z slab of a larger grid
    _emu(np.float64, "fill_hash", n, tmp_path, lo[0], lo[1], lo[2], ngl[0], ngl[1], nhalo, 4242)
    shp = (n[2] + 2 * nhalo, n[1] + 2 * nhalo, n[0] + 2 * nhalo)
    got = _read(tmp_path, "p_out", shp, np.float64)
    ref = np.zeros(shp)
    inner = (slice(nhalo, shp[0] - nhalo), slice(nhalo, shp[1] - nhalo), slice(nhalo, shp[2] - nhalo))
    ref[inner] = O.hash_field(ngl, 4242, lo=lo, n=n)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("walls", [False, True])
@pytest.mark.parametrize("ng,dt", SHAPES, ids=IDS)
def test_fused_fill_source(tmp_path, ng, dt, walls):
    """R2Fill / R2FillLine (cans_b200/csrc/r2_fill.cuh): what the loads of the fused forward x transform evaluate
    (`cansb200_solve_fillps`) == fillps followed by updt_rhs_b, on two z chunks of the slab, with a face-centred wall."""
    cf, cbc = ["c", "c", "f"], [["D", "D"], ["N", "N"], ["N", "D"]]
    cs = O.make_case(ng, [2.0, 1.0, 1.5], [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0, dtype=dt)
    f = _fields(ng, dt, tmp_path, "uvw")
    cs["dzfi"].tofile(os.path.join(tmp_path, "dzfi.bin"))
    dti = dt(1.0 / 0.37)
    ref = np.full(f["u"].shape, 3.25, dtype=dt)
    O.fillps(ng, cs["dli"], cs["dzfi"], dti, f["u"], f["v"], f["w"], ref)
    idx, vals = [0] * 6, [0.0] * 6
    if walls:
        rh = [[0.3, -0.7], [0.11, 0.05], [-0.4, 0.9]]
        O.updt_rhs_b(cf, cbc, ng, rh[0], rh[1], rh[2], ref)
        q = [1 if (cf[d] == "f" and cbc[d][1] == "D") else 0 for d in range(3)]
        idx = [v for d in range(3) for v in (1, ng[d] - q[d])]
        vals = [float(dt(rh[d][s])) for d in range(3) for s in range(2)]
    _emu(dt, "fill_source", ng, tmp_path, cs["dli"][0], cs["dli"][1], dti, *idx, *vals, max(1, ng[2] // 3))
    assert np.array_equal(_read(tmp_path, "p_out", ref.shape, dt), ref)
