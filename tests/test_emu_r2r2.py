"""The hot-path transform kernels' own source (the two-for-one register kernels of cans_b200/csrc/r2r2.cuh), compiled by g++
and run on the CPU with every CUDA thread of a CTA as a host thread and real barriers (tests/emu/emu_r2r2.cpp), against the
oracle's r2r definitions (FFTW's, src/fft.f90:247-258): EVERY plan instantiated in r2r2_inst.cuh -- all lengths 64 .. 2048,
all tuning variants, x (contiguous) and y (strided) mode, the predicated and the unpredicated y kernels, odd line counts,
FP64 and FP32 -- plus the split spectral order of the deduplicated pivot cache, the SPLIT kernels that store / load their
rows through a row table (the exchange of the distributed solve) and the forward x transform with the fused fillps +
updt_rhs_b source (`cansb200_solve_fillps`)."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import cans_oracle as O

BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")
KINDS = [0, 1, 5, 4, 9, 8]   # R2HC HC2R REDFT10 REDFT01 RODFT10 RODFT01
FWD = [0, 5, 9]


def _plans(macro):
    """(n, variant) of every plan in one X-macro table of r2r2_inst.cuh"""
    src = open(os.path.join(os.path.dirname(BUILD), "..", "cans_b200", "csrc", "r2r2_inst.cuh")).read()
    lines = src[src.index("#define " + macro):].split("\n")
    body = []
    for ln in lines:                      # the macro body: continuation lines up to the first one without a backslash
        body.append(ln)
        if not ln.rstrip().endswith("\\"):
            break
    return sorted({(int(a), int(b)) for a, b in re.findall(r"X\((\d+),\s*(\d+),", "\n".join(body))})


X_PLANS, Y_PLANS, Y32_PLANS = _plans("CB_R2_X_CONFIGS"), _plans("CB_R2_Y_CONFIGS"), _plans("CB_R2_Y32_CONFIGS")


def _emu(dt, mode, n, var, shape, d, *extra):
    exe = os.path.join(BUILD, "emu_r2r2_f32" if dt == np.float32 else "emu_r2r2_f64")
    if not os.path.exists(exe):   # built by __graft_entry__.build() (conftest's session fixture)
        pytest.skip("the r2r2 emulator was not built (g++ or the CUDA headers are missing)")
    nz, ny, nx = shape
    r = subprocess.run([exe, "f32" if dt == np.float32 else "f64", mode, str(n), str(var), str(nx), str(ny), str(nz), str(d)]
                       + [str(e) for e in extra], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout + r.stderr)


def _check(dt, mode, n, var, shape, d, kinds=KINDS, ref_of=None, tol=None):
    axis = 2 if mode == "x" else 1
    x = np.fromfile(os.path.join(d, "arr.bin"), dtype=dt).reshape(shape).astype(np.float64)
    tol = tol or ((3e-15 if dt == np.float64 else 2e-6) * max(1.0, np.log2(n)))
    for k in kinds:
        got = np.fromfile(os.path.join(d, f"out_{k}.bin"), dtype=dt).reshape(shape)
        ref = ref_of(x, k) if ref_of else O.r2r_1d(x, k, axis=axis)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < tol, f"{mode} n={n} var={var} kind={k}: {err:.2e}"


def _input(dt, shape, d, seed):
    x = np.random.default_rng(seed).uniform(-1, 1, shape).astype(dt)
    x.tofile(os.path.join(d, "arr.bin"))
    return x


@pytest.mark.parametrize("n,var", X_PLANS)
def test_x_plans(tmp_path, n, var):
    """contiguous lines; 5 lines = an odd count: the last pair has no second line"""
    shape = (1, 5, n)
    _input(np.float64, shape, tmp_path, n + var)
    _emu(np.float64, "x", n, var, shape, tmp_path)
    _check(np.float64, "x", n, var, shape, tmp_path)


@pytest.mark.parametrize("n,var,nx", [(n, v, 20 if n < 1024 else 12) for n, v in Y_PLANS] + [(n, v, 32) for n, v in Y_PLANS if v in (0, 3) and n < 1024])
def test_y_plans(tmp_path, n, var, nx):
    """strided lines; 20 columns = a ragged last tile (predicated kernels), 32 = whole tiles (unpredicated kernels)"""
    shape = (1, n, nx)
    _input(np.float64, shape, tmp_path, n + var + nx)
    _emu(np.float64, "y", n, var, shape, tmp_path)
    _check(np.float64, "y", n, var, shape, tmp_path)


@pytest.mark.parametrize("n,var", Y32_PLANS)
def test_y_plans_fp32(tmp_path, n, var):
    shape = (1, n, 40 if n < 1024 else 24)
    _input(np.float32, shape, tmp_path, n + var)
    _emu(np.float32, "y", n, var, shape, tmp_path)
    _check(np.float32, "y", n, var, shape, tmp_path)


@pytest.mark.parametrize("n", sorted({n for n, v in X_PLANS if v == 0}))
def test_x_plans_fp32(tmp_path, n):
    shape = (1, 3, n)
    _input(np.float32, shape, tmp_path, n)
    _emu(np.float32, "x", n, 0, shape, tmp_path)
    _check(np.float32, "x", n, 0, shape, tmp_path)


@pytest.mark.parametrize("n,var", [p for p in X_PLANS if p[1] in (0, 3)])
def test_x_split_order(tmp_path, n, var):
    """CB_R2_XSPLIT: inside a solve with the deduplicated pivot cache a periodic x direction keeps its spectrum as
    (r0 .. r[n/2-1] | r[n/2], i1 .. i[n/2-1]) instead of FFTW's halfcomplex order; the other kinds are unaffected"""
    shape = (1, 4, n)
    _input(np.float64, shape, tmp_path, 7 * n + var)
    _emu(np.float64, "x", n, var, shape, tmp_path, "xsplit")
    hc_of = np.array([s if 2 * s <= n else n - (s - n // 2) for s in range(n)])

    def ref_of(x, k):
        if k == 0:                                   # forward: halfcomplex result read in split order
            return O.r2r_1d(x, 0, axis=2)[..., hc_of]
        if k == 1:                                   # backward: the input is a spectrum in split order
            hc = np.empty_like(x)
            hc[..., hc_of] = x
            return O.r2r_1d(hc, 1, axis=2)
        return O.r2r_1d(x, k, axis=2)
    _check(np.float64, "x", n, var, shape, tmp_path, ref_of=ref_of)


@pytest.mark.parametrize("n,var", [p for p in Y_PLANS if p[1] in (0, 3)])   # the plans run_r2r picks by default
def test_y_plans_through_a_row_table(tmp_path, n, var):
    """SPLIT kernels (the exchange of the distributed solve): forward kinds store result row j of plane g at
    table[j].ptr + g * table[j].gs, backward kinds load their input rows from there.  The far side here is laid out like the
    way-back buffer, [row][plane][x], with the rows dealt out in a scrambled order."""
    nx, nz = (24 if n < 1024 else 8), 2
    shape = (nz, n, nx)
    _input(np.float64, shape, tmp_path, 11 * n + var)
    perm = np.random.default_rng(n).permutation(n)
    rows = np.concatenate([[n * nz * nx], perm * nz * nx, np.full(n, nx)]).astype(np.int64)
    rows.tofile(os.path.join(tmp_path, "rows.bin"))
    _emu(np.float64, "y", n, var, shape, tmp_path, "rows")
    _check(np.float64, "y", n, var, shape, tmp_path)


@pytest.mark.parametrize("walls", [False, True])
@pytest.mark.parametrize("n", sorted({n for n, v in X_PLANS if v == 0}))
def test_forward_x_with_the_fused_fillps_source(tmp_path, n, walls):
    """`cansb200_solve_fillps`: the forward x kernels instantiated on R2ArgsFill evaluate fillps (+ the wall terms of
    updt_rhs_b) in their loads == the transform of the oracle's fillps -> updt_rhs_b field"""
    ng, dt = [n, 3, 3], np.float64
    cf, cbc = ["c", "c", "c"], [["D", "D"], ["N", "N"], ["N", "D"]]
    cs = O.make_case(ng, [2.0, 1.0, 1.5], [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0, dtype=dt)
    hg = [ng[0] + 2, ng[1] + 2, ng[2] + 2]
    uvw = [(0.5 * O.hash_field(hg, 90 + s)).astype(dt) for s in range(3)]
    for nm, a in zip("uvw", uvw):
        a.tofile(os.path.join(tmp_path, nm + ".bin"))
    cs["dzfi"].tofile(os.path.join(tmp_path, "dzfi.bin"))
    dti = 1.0 / 0.37
    p = np.zeros(uvw[0].shape)
    O.fillps(ng, cs["dli"], cs["dzfi"], dti, uvw[0], uvw[1], uvw[2], p)
    idx, vals = [0] * 6, [0.0] * 6
    if walls:
        rh = [[0.3, -0.7], [0.11, 0.05], [-0.4, 0.9]]
        O.updt_rhs_b(cf, cbc, ng, rh[0], rh[1], rh[2], p)
        idx = [v for d in range(3) for v in (1, ng[d])]
        vals = [rh[d][s] for d in range(3) for s in range(2)]
    rhs = np.ascontiguousarray(p[1:-1, 1:-1, 1:-1])
    rhs.tofile(os.path.join(tmp_path, "arr.bin"))      # what _check transforms with the oracle
    shape = (ng[2], ng[1], ng[0])
    _emu(dt, "x", n, 0, shape, tmp_path, "fill", repr(float(cs["dli"][0])), repr(float(cs["dli"][1])), repr(dti), *idx,
         *[repr(v) for v in vals])
    _check(dt, "x", n, 0, shape, tmp_path, kinds=FWD)
