"""GPU: the reference's only golden vector replayed with the CUDA solver as the Poisson solve.

Same loop as tests/test_ldc_golden.py (oracle/ldc_replay.py restates the explicit CaNS time step), but
every one of the 4500 pressure solves goes through `cansb200_solve` (C ABI, device pointers).  The result is
held to the reference's own bar (tests/lid_driven_cavity/test.py:8: rtol = 1e-7) AND to the oracle replay."""
import importlib
import os

import numpy as np
import pytest

from oracle import cans_oracle as O
from oracle import ldc_replay as L

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ldc_re1000_ref.txt")


def _cuda_solve_factory(cb, cfg):
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
    return solve, (ctx, sd)


def test_cuda_solver_reproduces_the_reference_ldc_vector():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    cb = importlib.import_module("cans_b200")
    solve, keep = _cuda_solve_factory(cb, L.LDC)
    ref = np.loadtxt(GOLD)
    out, st = L.run_ldc(solve=solve, return_state=True)
    np.testing.assert_allclose(out, ref[:, 1], rtol=1e-7, atol=0)
    assert st["divmax"] < O.small(np.float64)
    # and against the oracle's own replay of the first 150 steps, much tighter
    short_gpu = L.run_ldc(solve=solve, nstep=150)
    short_cpu = L.run_ldc(nstep=150)
    np.testing.assert_allclose(short_gpu, short_cpu, rtol=1e-10, atol=1e-14)
