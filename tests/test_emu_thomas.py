"""The reference-order tridiagonal kernels' own source (`thomas_factor_kernel` + `thomas_seq_kernel` of
cans_b200/csrc/thomas_kernels.cuh, thomas_variant = 0), compiled by g++ and run one emulated thread per column
(tests/emu/emu_thomas.cpp), against the oracle's `gaussel` (src/solver.f90:114-307): **bit-identical**, including the
singular-pivot pin, the periodic closure, the lambda-less variant of `solver_gaussel_z` and the deduplicated pivot cache.
On the GPU the pipelined kernels are then held to this variant (tests/test_gpu_parity.py::test_gaussel_stage)."""
import os
import subprocess

import numpy as np
import pytest

import cases
from oracle import cans_oracle as O

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "emu_thomas")


def _emu(dt, nx, ny, nz, n_rows, periodic, nopin, dedx, dedy, norm, d):
    if not os.path.exists(EMU):   # built by __graft_entry__.build() (conftest's session fixture)
        pytest.skip("tests/_build/emu_thomas was not built (g++ or the CUDA headers are missing)")
    r = subprocess.run([EMU, "f32" if dt == np.float32 else "f64", str(nx), str(ny), str(nz), str(n_rows), str(int(periodic)),
                        str(int(nopin)), str(int(dedx)), str(int(dedy)), repr(float(norm)), str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["C1_ldc_2x64x64", "C2s_triperiodic", "C3s_channel", "C4s_duct", "periodic_z_odd", "tiny_z",
                                  "odd_sizes", "dirichlet_xyz", "helm_w_face_z", "helm_w_face_z_nn", "fp32_channel", "fp32_ldc"])
def test_reference_order_kernels_are_bit_identical_to_gaussel(tmp_path, name):
    cs = cases.build_case(name)
    ng, dt = cs["ng"], cs["dtype"]
    rng = np.random.default_rng(11)
    pz = rng.uniform(-1, 1, (ng[2], ng[1], ng[0])).astype(dt)
    q = 1 if (cs["c_or_f"][2] == "f" and cs["cbc"][2][1] == "D") else 0
    per = cs["cbc"][2] == cases.P
    ref = pz.copy()
    O.gaussel(ng[2] - q, cs["a"], cs["b"], cs["c"], per, cs["normfft"], ref, cs["lambdaxy"])
    for nm, arr in (("p", pz), ("lam", cs["lambdaxy"]), ("a", cs["a"]), ("b", cs["b"]), ("c", cs["c"])):
        np.ascontiguousarray(arr).tofile(os.path.join(tmp_path, nm + ".bin"))
    _emu(dt, ng[0], ng[1], ng[2], ng[2] - q, per, 0, 0, 0, cs["normfft"], tmp_path)
    got = np.fromfile(os.path.join(tmp_path, "p_out.bin"), dtype=dt).reshape(pz.shape)
    assert np.array_equal(got, ref, equal_nan=True), f"{name}: rel L2 {cases.rel_l2(got, ref):.2e}"


@pytest.mark.parametrize("name", ["C3s_channel", "periodic_z_odd", "helm_w_face_z"])
def test_lambda_less_variant(tmp_path, name):
    """solver_gaussel_z (src/solver.f90:547-616): gaussel without lambdaxy, i.e. no singular-pivot pin and no tolerance
    test on the periodic closure (`nopin`), on an all-zero lambda."""
    cs = cases.build_case(name)
    ng, dt = cs["ng"], cs["dtype"]
    rng = np.random.default_rng(12)
    pz = rng.uniform(-1, 1, (ng[2], ng[1], ng[0])).astype(dt)
    q = 1 if (cs["c_or_f"][2] == "f" and cs["cbc"][2][1] == "D") else 0
    per = cs["cbc"][2] == cases.P
    bb = (cs["b"] + dt(1.0) / dt(cases.ALPHA)).astype(dt)
    ref = pz.copy()
    O.gaussel(ng[2] - q, cs["a"], bb, cs["c"], per, 1.0 / cases.ALPHA, ref, None)
    for nm, arr in (("p", pz), ("lam", np.zeros((ng[1], ng[0]), dt)), ("a", cs["a"]), ("b", bb), ("c", cs["c"])):
        np.ascontiguousarray(arr).tofile(os.path.join(tmp_path, nm + ".bin"))
    _emu(dt, ng[0], ng[1], ng[2], ng[2] - q, per, 1, 0, 0, 1.0 / cases.ALPHA, tmp_path)
    got = np.fromfile(os.path.join(tmp_path, "p_out.bin"), dtype=dt).reshape(pz.shape)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dedx,dedy", [(1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("per", [False, True])
def test_deduplicated_pivot_cache(tmp_path, dedx, dedy, per):
    """One stored copy of the pivots of the columns that share an eigenvalue (x: split order, position i >= nx/2 + 16 uses
    i - nx/2; y: halfcomplex order, row j > ny/2 uses ny - j): with an exactly mirror-symmetric lambdaxy the result is the
    full cache's, bit for bit."""
    nx, ny, nz = 64, 10, 23
    cbc = [cases.P, cases.P, cases.P if per else cases.N]
    cs = O.make_case([nx, ny, nz], [6.0, 3.0, 2.0], cbc, gr=0.0 if per else 1.5)
    lam = cs["lambdaxy"].copy()                       # halfcomplex order in x and y
    hc_of = [s if 2 * s <= nx else nx - (s - nx // 2) for s in range(nx)]
    lam = lam[:, hc_of]                               # x in split order, as inside a solve with the deduplicated cache
    lam[:, nx // 2 + 16:] = lam[:, 16:nx // 2]        # exact mirror symmetry (initsolver's two halves agree to rounding only)
    for j in range(ny // 2 + 1, ny):
        lam[j] = lam[ny - j]
    rng = np.random.default_rng(13)
    pz = rng.uniform(-1, 1, (nz, ny, nx))
    ref = pz.copy()
    O.gaussel(nz, cs["a"], cs["b"], cs["c"], per, cs["normfft"], ref, lam)
    for nm, arr in (("p", pz), ("lam", lam), ("a", cs["a"]), ("b", cs["b"]), ("c", cs["c"])):
        np.ascontiguousarray(arr).tofile(os.path.join(tmp_path, nm + ".bin"))
    _emu(np.float64, nx, ny, nz, nz, per, 0, dedx, dedy, cs["normfft"], tmp_path)
    got = np.fromfile(os.path.join(tmp_path, "p_out.bin"), dtype=np.float64).reshape(pz.shape)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name,nsplit", [("C3s_channel", 2), ("C3s_channel", 4), ("C4s_duct", 3), ("periodic_z_odd", 3),
                                         ("C2s_triperiodic", 4), ("helm_w_face_z", 2), ("fp32_channel", 4), ("C1_ldc_2x64x64", 1)])
def test_distributed_tdma_kernels_are_bit_identical_to_gaussel_dtdma(tmp_path, name, nsplit):
    """The four kernels of dtdma_kernels.cuh, driven as `cansb200_gaussel_dtdma` drives them, against the oracle's
    `gaussel_dtdma` (src/solver.f90:309-517): slab-wise elimination, reduced 2-rows-per-rank system (periodic closure
    included), update -- even and uneven z splits, with and without lambdaxy."""
    from cans_b200.decomp import split_starts
    if not os.path.exists(EMU):
        pytest.skip("tests/_build/emu_thomas was not built")
    cs = cases.build_case(name)
    ng, dt = cs["ng"], cs["dtype"]
    helm = name in cases.HELMHOLTZ
    a, c = cs["a"], cs["c"]
    b = (cs["b"] + dt(1.0 / cases.ALPHA if helm else 0.0)).astype(dt)
    lam = (cs["lambdaxy"] - dt(0.0 if helm or not cases.is_singular(cs) else 0.37)).astype(dt)   # regular columns
    per = cs["cbc"][2] == cases.P
    q = 1 if (cs["c_or_f"][2] == "f" and cs["cbc"][2][1] == "D") else 0
    n = ng[2] - q
    starts = split_starts(ng[2], nsplit)
    rng = np.random.default_rng(29)
    pz = rng.uniform(-1, 1, (ng[2], ng[1], ng[0])).astype(dt)
    norm = 0.61
    for use_lam, bb in ((True, b), (False, (b - dt(3.0)).astype(dt))):
        ref = pz.copy()
        O.gaussel_dtdma(starts, n, a, bb, c, per, norm, ref, lam if use_lam else None)
        for nm, arr in (("p", pz), ("lam", lam), ("a", a), ("b", bb), ("c", c)):
            np.ascontiguousarray(arr).tofile(os.path.join(tmp_path, nm + ".bin"))
        r = subprocess.run([EMU, "dtdma", "f32" if dt == np.float32 else "f64", str(ng[0]), str(ng[1]), str(ng[2]), str(n),
                            str(int(per)), str(int(use_lam)), repr(norm), str(tmp_path), str(nsplit)] + [str(v) for v in starts],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        got = np.fromfile(os.path.join(tmp_path, "p_out.bin"), dtype=dt).reshape(pz.shape)
        assert np.array_equal(got, ref), f"{name} lam={use_lam}: rel L2 {cases.rel_l2(got, ref):.2e}"
