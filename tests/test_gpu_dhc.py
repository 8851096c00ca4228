"""GPU, opt-in (CANSB200_LONG_TESTS=1; written after the round's GPU budget was spent, so it has not run on a B200 yet and is
skipped by default): the reference's differentially-heated-cavity case replayed with the CUDA solver as the Poisson solve.

Same loop as tests/test_dhc_golden.py (oracle/dhc_replay.py restates the CaNS time step with one scalar and Boussinesq
buoyancy); every pressure solve goes through `cansb200_solve` (C ABI, device pointers).  On the solver path this case has
REDFT10 / REDFT01 along x (n = 128) -- the lid-driven cavity has them along y only.  The full 10 000 steps (the reference's
Nusselt number) take ten minutes of host time; here the first steps are held to the oracle's own replay, tightly."""
import importlib

import os

import numpy as np
import pytest

from oracle import cans_oracle as O
from oracle import dhc_replay as D

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CANSB200_LONG_TESTS") != "1", reason="opt-in: set CANSB200_LONG_TESTS=1")]
torch = pytest.importorskip("torch")


def test_cuda_solver_follows_the_oracle_on_the_dhc_case():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    cb = importlib.import_module("cans_b200")
    cfg = D.DHC
    dev = torch.device("cuda:0")
    ng, l = cfg["ng"], cfg["l"]
    cs = O.make_case(ng, l, cfg["cbcpre"], gr=cfg["gr"], gtype=cfg["gtype"], bc=cfg["bcpre"])
    ctx = cb.Context(ng)
    sd = cb.initsolver(ctx, ng, cs["dli"], cs["dzci"], cs["dzfi"], cfg["cbcpre"], cfg["bcpre"], ["c"] * 3, device=dev)
    pd = torch.empty((ng[2] + 2, ng[1] + 2, ng[0] + 2), dtype=torch.float64, device=dev)

    def solve(pp):
        pd.copy_(torch.from_numpy(pp))
        cb.solver(ng, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cfg["cbcpre"], ["c"] * 3, pd)
        pp[...] = pd.cpu().numpy()

    nstep = 30
    nu_g, st_g = D.run_dhc(solve=solve, nstep=nstep, return_state=True)
    nu_c, st_c = D.run_dhc(nstep=nstep, return_state=True)
    assert st_g["divmax"] < O.small(np.float64)
    np.testing.assert_allclose(nu_g, nu_c, rtol=1e-11)
    for k in ("s", "u", "w"):
        np.testing.assert_allclose(st_g[k], st_c[k], rtol=1e-9, atol=1e-13)
