"""GPU: the steps either side of the solver path against the oracle, kernel by kernel:
fillps (src/fillps.f90:38-50), correc (src/correc.f90:33-59), chkdiv (src/chkdiv.f90:35-50) -- both launch geometries
(CANSB200_CTX_AUX_3D: 3-D grid without index divisions, and the flat-index kernels it falls back to)."""
import importlib

import numpy as np
import pytest

from oracle import cans_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

# (ng, dtype): the lid-driven-cavity shape (nx = 2), ragged extents that do not fill a CTA, a row longer than one CTA
SHAPES = [([2, 64, 64], np.float64), ([37, 5, 9], np.float64), ([300, 3, 4], np.float64), ([64, 48, 40], np.float64),
          ([33, 17, 6], np.float32)]


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    return importlib.import_module("cans_b200")


@pytest.mark.parametrize("geom3d", [True, False])
@pytest.mark.parametrize("ng,dt", SHAPES, ids=[f"{s[0][0]}x{s[0][1]}x{s[0][2]}_{np.dtype(s[1]).name}" for s in SHAPES])
def test_fillps_correc_chkdiv_match_oracle(cb, ng, dt, geom3d):
    S = importlib.import_module("cans_b200.solver")
    dev = torch.device("cuda:0")
    l = [2.0, 1.0, 1.5]
    cs = O.make_case(ng, l, [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0, dtype=dt)
    hg = [ng[0] + 2, ng[1] + 2, ng[2] + 2]
    u, v, w, p = ((0.5 * O.hash_field(hg, 40 + s)).astype(dt) for s in range(4))
    ctx = cb.Context(ng, is_fp32=dt == np.float32)
    ctx.set_aux_3d(geom3d)
    ud, vd, wd, pd = (torch.from_numpy(t.copy()).to(dev) for t in (u, v, w, p))
    dzfi, dzci = torch.from_numpy(cs["dzfi"]).to(dev), torch.from_numpy(cs["dzci"]).to(dev)
    eps = np.finfo(dt).eps
    dti, dtc = 1.0 / 0.37, 0.37
    # fillps: interior of p, halos untouched (the compiler may contract a*b+c: a few ulp of the largest term)
    pr = p.copy()
    O.fillps(ng, cs["dli"], cs["dzfi"], dti, u, v, w, pr)
    S.fillps(ctx, ng, cs["dli"], dzfi, dti, ud, vd, wd, pd)
    got = pd.cpu().numpy()
    assert np.abs(got - pr).max() <= 16 * eps * np.abs(pr).max()
    halo = np.ones(pr.shape, bool)
    halo[1:-1, 1:-1, 1:-1] = False
    assert np.array_equal(got[halo], p[halo])
    # chkdiv of the same field
    tot_o, max_o = O.chkdiv(ng, l, cs["dli"], cs["dzfi"], u, v, w)
    tot_g, max_g = S.chkdiv(ctx, ng, l, cs["dli"], dzfi, ud, vd, wd)
    assert abs(max_g - max_o) <= 64 * eps * max_o
    assert abs(tot_g - tot_o) <= (1e-12 if dt == np.float64 else 1e-4) * tot_o
    # correc: every loop of the reference runs over its own haloed range
    ur, vr, wr = u.copy(), v.copy(), w.copy()
    O.correc(ng, cs["dli"], cs["dzci"], dtc, p, ur, vr, wr)
    pd.copy_(torch.from_numpy(p))
    S.correc(ctx, ng, cs["dli"], dzci, dtc, pd, ud, vd, wd)
    for g, r, name in ((ud, ur, "u"), (vd, vr, "v"), (wd, wr, "w")):
        gh = g.cpu().numpy()
        scale = np.abs(r).max()
        assert np.abs(gh - r).max() <= 8 * eps * scale * max(1.0, float(np.max(cs["dli"])) * dtc), name
    assert np.array_equal(pd.cpu().numpy(), p), "correc must not touch p"
    ctx.close()
