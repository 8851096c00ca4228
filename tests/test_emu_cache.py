"""The kernels that decide WHICH factorisation a solve uses, and the acceptance metric, from their own source on the CPU
(tests/emu/emu_cache.cpp: host-thread model with warp shuffles and atomics):

* `thomas_hash_kernel` + `thomas_select_kernel` (cans_b200/csrc/thomas_kernels.cuh): the pivot cache is keyed by two independent
  content hashes of (a, b, c, lambdaxy) computed on the device every solve, slots are reused least-recently-used.  This is what
  replaces the reference's `is_dtdma_update` / "the caller knows when b changed" (src/solve_helmholtz.f90:63-71): a wrong hit
  would silently solve with the wrong pivots.
* `chkdiv_kernel` (aux_kernels.cuh, src/chkdiv.f90:35-50) against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import cans_oracle as O

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "emu_cache")


def _run(*args):
    if not os.path.exists(EMU):   # built by __graft_entry__.build() (conftest's session fixture)
        pytest.skip("tests/_build/emu_cache was not built (g++ or the CUDA headers are missing)")
    r = subprocess.run([EMU] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout + r.stderr)
    return r.stdout


def _write_sets(d, sets):
    for s, (a, b, c, lam) in enumerate(sets):
        for nm, arr in (("a", a), ("b", b), ("c", c), ("lam", lam)):
            np.ascontiguousarray(arr, dtype=np.float64).tofile(os.path.join(d, f"set_{s}_{nm}.bin"))


def _cache(d, nx, ny, n, nslots, nsets, seq):
    out = _run("cache", nx, ny, n, nslots, nsets, d, *seq)
    rows = [ln.split() for ln in out.strip().splitlines()]
    return [dict(hit=int(r[0]), sel=int(r[1]), nfactor=int(r[2]), sym_bad=int(r[3]), key=int(r[4]), key2=int(r[5])) for r in rows]


def test_slots_are_reused_least_recently_used(tmp_path):
    """three Helmholtz coefficient sets (one per RK sub-step alpha) + the Poisson set on a two-slot and a three-slot cache"""
    nx, ny, n = 24, 10, 37
    cs = O.make_case([nx, ny, n], [2.0, 1.0, 1.5], [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0)
    sets = [(cs["a"], cs["b"] + sh, cs["c"], cs["lambdaxy"]) for sh in (0.0, -3.0, -7.5, -11.0)]
    _write_sets(tmp_path, sets)
    r = _cache(tmp_path, nx, ny, n, 2, 4, ["0:000", "1:000", "0:000", "2:000", "1:000", "1:000"])
    assert [x["hit"] for x in r] == [0, 0, 1, 0, 0, 1]
    assert [x["sel"] for x in r] == [0, 1, 0, 1, 0, 0]      # set 2 evicts set 1 (slot 1), then set 1 evicts set 0 (slot 0)
    assert [x["nfactor"] for x in r] == [1, 2, 2, 3, 4, 4]
    # three slots hold the three alphas of an RK3 step: after the first step every solve is a hit
    r = _cache(tmp_path, nx, ny, n, 3, 4, ["0:000", "1:000", "2:000"] * 3)
    assert [x["hit"] for x in r] == [0, 0, 0] + [1] * 6 and [x["sel"] for x in r] == [0, 1, 2] * 3
    assert len({(x["key"], x["key2"]) for x in r}) == 3


def test_hash_sees_every_element_its_position_and_the_solve_variant(tmp_path):
    nx, ny, n = 16, 6, 21
    cs = O.make_case([nx, ny, n], [2.0, 1.0, 1.5], [["P", "P"], ["P", "P"], ["N", "N"]], gr=1.5)
    base = (cs["a"].copy(), cs["b"].copy(), cs["c"].copy(), cs["lambdaxy"].copy())
    sets = [base]
    for which, idx in ((0, 3), (1, n - 1), (2, 0), (3, (ny - 1, nx - 1))):   # one ulp in a, b, c, lambdaxy
        s = [v.copy() for v in base]
        s[which][idx] = np.nextafter(s[which][idx], np.inf)
        sets.append(tuple(s))
    s = [v.copy() for v in base]
    s[1][[4, 5]] = s[1][[5, 4]]                                             # two entries of b swapped: same multiset
    sets.append(tuple(s))
    _write_sets(tmp_path, sets)
    seq = [f"{i}:000" for i in range(len(sets))] + ["0:100", "0:000"]       # ... and the lambda-less (nopin) variant of set 0
    r = _cache(tmp_path, nx, ny, n, 8, len(sets), seq)
    assert [x["hit"] for x in r] == [0] * (len(sets) + 1) + [1]
    assert len({x["key"] for x in r[:-1]}) == len(sets) + 1 and len({x["key2"] for x in r[:-2]}) == len(sets)
    assert r[-1]["sel"] == 0


def test_symmetry_check_of_the_deduplicated_cache(tmp_path):
    """dx / dy: the cache keeps one copy of the pivots of mirror columns, so the hash kernel verifies lambdaxy's mirror
    symmetry on every solve (x in split order: position i >= nx/2 + 16 mirrors i - nx/2; y: row j > ny/2 mirrors ny - j)"""
    nx, ny, n = 64, 8, 9
    cs = O.make_case([nx, ny, n], [6.0, 3.0, 2.0], [["P", "P"], ["P", "P"], ["N", "N"]], gr=1.5)
    hc_of = [s if 2 * s <= nx else nx - (s - nx // 2) for s in range(nx)]
    lam = cs["lambdaxy"][:, hc_of]                      # as lambda_unpack hands it to the solve: x in split order
    bad_x, bad_y = lam.copy(), lam.copy()
    bad_x[2, nx // 2 + 20] *= 1.0 + 1e-8                # breaks the x mirror (column 52 against column 20) ...
    bad_x[ny - 2, nx // 2 + 20] = bad_x[2, nx // 2 + 20]   # ... and keeps the y mirror (row 6 against row 2)
    bad_y[ny - 1, 3] *= 1.0 + 1e-8                      # row 7 against row 1; column 3 has no stored x mirror
    _write_sets(tmp_path, [(cs["a"], cs["b"], cs["c"], m) for m in (lam, bad_x, bad_y)])
    for seq, want in ((["0:011"], 0), (["1:010"], 1), (["1:001"], 0), (["2:001"], 1), (["2:010"], 0), (["1:000"], 0)):
        r = _cache(tmp_path, nx, ny, n, 2, 3, seq)
        assert r[0]["sym_bad"] == want, (seq, r)


@pytest.mark.parametrize("ng,dt", [([37, 5, 9], np.float64), ([64, 12, 10], np.float64), ([33, 17, 6], np.float32)])
def test_chkdiv_kernel(tmp_path, ng, dt):
    l = [2.0, 1.0, 1.5]
    cs = O.make_case(ng, l, [["P", "P"], ["N", "N"], ["D", "D"]], gr=1.0, dtype=dt)
    hg = [ng[0] + 2, ng[1] + 2, ng[2] + 2]
    uvw = [(0.5 * O.hash_field(hg, 50 + s)).astype(dt) for s in range(3)]
    for nm, a in zip("uvw", uvw):
        a.tofile(os.path.join(tmp_path, nm + ".bin"))
    cs["dzfi"].tofile(os.path.join(tmp_path, "dzfi.bin"))
    out = _run("chkdiv", "f32" if dt == np.float32 else "f64", *ng, repr(float(cs["dli"][0])), repr(float(cs["dli"][1])), tmp_path)
    tot, mx = (float(v) for v in out.split())
    tot_o, max_o = O.chkdiv(ng, l, cs["dli"], cs["dzfi"], *uvw)
    assert mx == max_o                                   # same expression, same order: the maximum is bit-exact
    assert tot / (l[0] * l[1] * l[2]) == pytest.approx(tot_o, rel=1e-13 if dt == np.float64 else 1e-5)


R2R_KINDS = {"R2HC": 0, "HC2R": 1, "REDFT00": 3, "REDFT01": 4, "REDFT10": 5, "REDFT11": 6,
             "RODFT00": 7, "RODFT01": 8, "RODFT10": 9, "RODFT11": 10}


@pytest.mark.parametrize("kname", sorted(R2R_KINDS))
@pytest.mark.parametrize("n,axis", [(2, 0), (3, 1), (7, 0), (17, 1), (30, 0), (31, 1)])
def test_direct_transform_kernel(tmp_path, kname, n, axis):
    """r2r_direct_kernel: every FFTW kind the reference can plan (src/fft.f90:260-313), odd / prime lengths, both axes, and a
    line longer than the transform (the face-centred Dirichlet case transforms n - 1 points and copies the last one)"""
    kind = R2R_KINDS[kname]
    if kname == "REDFT00" and n < 2:
        pytest.skip("REDFT00 needs n >= 2")
    ll = n + (1 if n % 2 else 0)                      # odd n: one untransformed tail point
    shape = (2, 3, ll) if axis == 0 else (2, ll, 5)
    x = np.random.default_rng(n * 31 + kind).uniform(-1, 1, shape)
    x.tofile(os.path.join(tmp_path, "arr.bin"))
    _run("direct", "f64", kind, n, axis, shape[2], shape[1], shape[0], tmp_path)
    got = np.fromfile(os.path.join(tmp_path, "out.bin"), dtype=np.float64).reshape(shape)
    ref = x.copy()
    sl = [slice(None)] * 3
    sl[2 - axis] = slice(0, n)
    ref[tuple(sl)] = O.r2r_1d(np.ascontiguousarray(x[tuple(sl)]), kind, axis=2 - axis)
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-14 * max(1.0, np.log2(n))
