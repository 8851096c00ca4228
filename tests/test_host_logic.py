"""CPU tests of the host-side logic added in round 2 (no GPU, no compute calls into the CUDA library):

  * the `_OPENACC` eigenvalue order of `initsolver` (src/initsolver.f90:98-117) and the index maps the kernels use for it
    (`pack_index`, `split_to_hc` in cans_b200/csrc/aux_kernels.cuh, restated here and checked as permutations);
  * `updt_rhs_b` with `is_bound` on z slabs against the single-rank restatement of src/bound.f90:514-598;
  * both arms of bench.py describe the same `config`, and the CPU arm neither imports nor loads the product.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from oracle import cans_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack_index(h, n):
    """twin of cb::pack_index: halfcomplex position -> position in the _OPENACC packed order"""
    if h == 0:
        return 0
    if 2 * h <= n:
        return 1 if 2 * h == n else 2 * h
    k = n - h
    return 2 * k + 1 if 2 * k + 1 < n else 1


def split_to_hc(pos, n):
    """twin of cb::split_to_hc: position in the split order (r0..r[n/2-1] | r[n/2], i1..) -> halfcomplex position"""
    return pos if 2 * pos <= n else n - (pos - n // 2)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 9, 16, 17, 64, 96, 250, 1024])
def test_packed_eigenvalue_order(n):
    """lambda_halfcomplex[h] == lambda_packed[pack_index(h)] for the reference's own `iswap` permutation, and pack_index is a
    bijection; split order is a bijection whose two halves pair the real and the imaginary part of a mode."""
    from cans_b200.solver import eigenvalues
    hc = eigenvalues(n, "PP", "c")
    pk = eigenvalues(n, "PP", "c", openacc_order=True)
    q = [pack_index(h, n) for h in range(n)]
    assert sorted(q) == list(range(n))
    # position q(h) of the packed array holds the eigenvalue of the same mode (bitwise for the real parts; the imaginary
    # part of mode k sits at halfcomplex n - k and at packed 2k + 1, both computed from frequency n - k)
    for h in range(n):
        assert pk[q[h]] == pytest.approx(hc[h], rel=1e-12, abs=1e-300)
    # the reference's description of the format: (r0, r[n/2], r1, i1, r2, i2, ...)
    lam = lambda k: -2.0 * (1.0 - np.cos(2.0 * np.pi * k / n))
    if n % 2 == 0 and n >= 4:
        assert pk[0] == pytest.approx(lam(0), abs=1e-15) and pk[1] == pytest.approx(lam(n // 2))
        for k in range(1, n // 2):
            assert pk[2 * k] == pytest.approx(lam(k)) and pk[2 * k + 1] == pytest.approx(lam(k))
    if n % 2 == 0 and n >= 4:
        s = [split_to_hc(p, n) for p in range(n)]
        assert sorted(s) == list(range(n))
        for p in range(1, n // 2):   # position p: real part of mode p; position n/2 + p: its imaginary part
            assert hc[s[p]] == pytest.approx(hc[s[n // 2 + p]], rel=1e-10)


@pytest.mark.parametrize("nranks", [1, 2, 3])
@pytest.mark.parametrize("cf,cbc", [(["c", "c", "c"], [cases.D, cases.D, cases.D]), (["c", "c", "f"], [cases.P, cases.N, cases.D]),
                                    (["f", "f", "c"], [cases.D, cases.D, cases.N])])
def test_updt_rhs_b_is_bound_on_slabs(nranks, cf, cbc):
    """The z contributions (and the face-centred Dirichlet shortening q) belong to the ranks that own the walls only."""
    from cans_b200.decomp import SlabDecomp
    from cans_b200.solver import updt_rhs_b
    ng = [7, 6, 11]
    rng = np.random.default_rng(0)
    p = rng.uniform(-1, 1, (ng[2] + 2, ng[1] + 2, ng[0] + 2))
    rx, ry, rz = [0.3, -1.7], [2.5, 0.125], [-0.75, 1.1]
    ref = p.copy()
    O.updt_rhs_b(cf, cbc, ng, rx, ry, rz, ref, cases.ALPHA)
    out = []
    for r in range(nranks):
        dec = SlabDecomp(ng, nranks, r)
        z0, z1 = dec.z_range()
        slab = np.ascontiguousarray(p[z0:z1 + 2])
        is_bound = [[True, True], [True, True], [z0 == 0, z1 == ng[2]]]
        updt_rhs_b(cf, cbc, dec.n, is_bound, rx, ry, rz, slab, cases.ALPHA)
        out.append(slab[1:-1])
    assert np.array_equal(np.concatenate(out, axis=0), ref[1:-1])


def test_bench_arms_share_config_and_cpu_arm_is_product_free():
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--workload','T_smoke_128x64x96','--steps','2','--warmup','1'];"
            "import runpy; runpy.run_path('bench.py', run_name='__main__');"
            "assert not any(m.startswith('cans_b200') for m in sys.modules), 'the CPU arm imported the product package';"
            "import ctypes; maps=open('/proc/self/maps').read(); assert 'libcans_b200' not in maps, 'the CPU arm loaded the product library'")
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports: the arm must override it
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and line["steps"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config("T_smoke_128x64x96", 1)   # what the product arm prints as `config`
    assert line["ms_per_step"] > 0 and line["value"] == pytest.approx(line["ms_per_step"] * 1e6 / (128 * 64 * 96))
