"""The five stages of `solver` (src/solver.f90:17-112: fft x, fft y, gaussel, ifft y, ifft x) composed on the CPU from the CUDA
kernels' own source -- the two-for-one transform kernels (tests/emu/emu_r2r2.cpp) and the reference-order tridiagonal kernels
(tests/emu/emu_thomas.cpp), chained the way `solve_impl` (capi.cu) chains them -- against the oracle's `solver`, at the
north-star tolerance of 1e-12.  Also the variant with the deduplicated pivot cache: x kept in split spectral order between the
two x transforms, eigenvalues permuted to match (`lambda_unpack_kernel`), mirror columns sharing their pivots."""
import os
import subprocess

import numpy as np
import pytest

import cases
from cans_b200.solver import find_fft
from oracle import cans_oracle as O

BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")


def _run(exe, *args):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.skip(f"tests/_build/{exe} was not built (g++ or the CUDA headers are missing)")
    r = subprocess.run([path] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (exe, r.returncode, r.stdout + r.stderr)


def _r2r(d, mode, kind, field, *extra):
    nz, ny, nx = field.shape
    n = nx if mode == "x" else ny
    np.ascontiguousarray(field).tofile(os.path.join(d, "arr.bin"))
    _run("emu_r2r2_f64", "f64", mode, n, 0, nx, ny, nz, d, *extra)
    return np.fromfile(os.path.join(d, f"out_{kind}.bin"), dtype=np.float64).reshape(field.shape)


@pytest.mark.parametrize("dedup", [False, True])
@pytest.mark.parametrize("ng,cbc,gr", [([64, 64, 12], [cases.P, cases.P, cases.N], 2.0),      # the channel operator (C3), singular
                                       ([128, 64, 9], [cases.P, cases.N, cases.N], 1.5),       # the duct operator (C4)
                                       ([64, 128, 11], [cases.D, cases.N, cases.D], 1.0),      # Dirichlet walls in x and z
                                       ([64, 64, 10], [cases.P, cases.P, cases.P], 0.0)])      # triperiodic (C2): cyclic closure
def test_kernel_chain_reproduces_the_oracle_solver(tmp_path, ng, cbc, gr, dedup):
    if dedup and cbc[0] != cases.P:
        pytest.skip("the split order exists for a periodic x direction only")
    d = str(tmp_path)
    cs = O.make_case(ng, [6.0, 3.0, 2.0], cbc, gr=gr)
    p = cases.make_rhs(cs, seed=321)
    ref = cases.oracle_solve("chain", cs, p)
    kxf, kxb, _ = find_fft(cbc[0], "c")
    kyf, kyb, _ = find_fft(cbc[1], "c")
    per = cbc[2] == cases.P
    xs = ("xsplit",) if dedup else ()
    f = np.ascontiguousarray(p[1:-1, 1:-1, 1:-1])
    f = _r2r(d, "x", kxf, f, *xs)                 # forward x (split spectral order with the deduplicated cache)
    f = _r2r(d, "y", kyf, f)                      # forward y
    lam = cs["lambdaxy"]
    if dedup:                                     # eigenvalues in the order the solve keeps the spectrum in
        lam.tofile(os.path.join(d, "lam.bin"))
        _run("emu_aux", "f64", "lambda_unpack", ng[0], ng[1], 1, d, 0, 0, 1)
        lam = np.fromfile(os.path.join(d, "lam_out.bin"), dtype=np.float64).reshape(lam.shape)
    for nm, arr in (("p", f), ("lam", lam), ("a", cs["a"]), ("b", cs["b"]), ("c", cs["c"])):
        np.ascontiguousarray(arr).tofile(os.path.join(d, nm + ".bin"))
    dedy = int(dedup and cbc[1] == cases.P)
    _run("emu_thomas", "f64", ng[0], ng[1], ng[2], ng[2], int(per), 0, int(dedup), dedy, repr(float(cs["normfft"])), d)
    f = np.fromfile(os.path.join(d, "p_out.bin"), dtype=np.float64).reshape(f.shape)
    f = _r2r(d, "y", kyb, f)                      # backward y
    f = _r2r(d, "x", kxb, f, *xs)                 # backward x
    err = cases.parity_error(cs, f, ref[1:-1, 1:-1, 1:-1])
    assert err < 1e-12, f"{ng} {cbc} dedup={dedup}: rel L2 {err:.3e}"
