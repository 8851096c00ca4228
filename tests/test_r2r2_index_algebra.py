"""CPU check of the index algebra the two-for-one transform kernels rely on (cans_b200/csrc/r2r2.cuh).

The kernels address shared memory as `enc(per-thread part) (+ or ^) enc(compile-time part)` instead of
`enc(whole position)`.  That is only valid when, for every plan in the instantiation table
(cans_b200/csrc/r2r2_inst.cuh),
  * digit reversal of a power-of-two plan is a bit permutation, so rev(t + c) = rev(t) + rev(c) for a thread
    index t < TPL and a multiple c of TPL (R2Pair::slot_k / slot_m),
  * the XOR swizzle of the x layout is GF(2)-linear (R2Lay::enc / join),
  * the positions a thread touches in a stage split into a thread part and an (m, q) part with disjoint bits
    (power of two) or, in y mode, plainly additive parts whenever L and TPL divide one another,
  * the mirror frequency N - k of k = t + m TPL is (TPL - t) + (N - TPL - m TPL) for t > 0.
This test re-derives every slot with plain integer arithmetic and compares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _configs():
    txt = open(os.path.join(ROOT, "cans_b200", "csrc", "r2r2_inst.cuh")).read()
    out = []
    for table, ymode, f32 in (("CB_R2_X_CONFIGS", False, False), ("CB_R2_Y_CONFIGS", True, False), ("CB_R2_Y32_CONFIGS", True, True)):
        body = txt[txt.index("#define " + table):]
        body = body[:body.index("\n\n") if "\n\n" in body else len(body)]
        body = body.split("\n// ")[0]
        for m in re.finditer(r"X\((\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\)", body):
            n, var, tpl, g, minb, r0, r1, r2, r3 = (int(v) for v in m.groups())
            out.append(dict(table=table, ymode=ymode, f32=f32, N=n, var=var, TPL=tpl, G=g, R=[r for r in (r0, r1, r2, r3) if r > 1]))
    return out


CONFIGS = _configs()


def rev(cfg, k):
    """R2Cfg::rev: slot of frequency k after the DIF stages."""
    n = cfg["N"]
    rad = cfg["R"] + [1] * (4 - len(cfg["R"]))
    pos, prod = 0, 1
    for r in rad[:3]:
        prod *= r
        pos += (k % r) * (n // prod)
        k //= r
    return pos + k


def enc(cfg, pos, esz=8):
    if cfg["ymode"]:
        return pos
    sw = 3 if esz == 8 else 4
    return pos ^ (((pos >> sw) ^ (pos >> (2 * sw))) & ((1 << sw) - 1))


def join(cfg, ea, eb):
    return ea + eb if cfg["ymode"] else ea ^ eb


def test_table_is_parsed():
    assert len(CONFIGS) >= 30
    assert any(c["N"] == 1024 and not c["ymode"] for c in CONFIGS)


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"{c['table'][6:]}-{c['N']}v{c['var']}")
def test_plan_constraints(cfg):
    n, tpl, rad = cfg["N"], cfg["TPL"], cfg["R"]
    prod = 1
    for r in rad:
        prod *= r
    assert prod == n
    assert tpl & (tpl - 1) == 0, "threads per transform must be a power of two"
    e = n // tpl
    assert n % tpl == 0 and all(e % r == 0 for r in rad), "every radix must divide E"
    assert n % 2 == 0 and (n // 2) % tpl == 0, "pair pass needs TPL | N/2"
    assert sorted(rev(cfg, k) for k in range(n)) == list(range(n)), "digit reversal must be a permutation"
    threads = tpl * cfg["G"]
    assert threads <= 1024
    esz = 4 if cfg["f32"] else 8
    smem = n * cfg["G"] * 2 * esz
    assert smem <= 227 * 1024


@pytest.mark.parametrize("cfg", [c for c in CONFIGS if c["N"] & (c["N"] - 1) == 0], ids=lambda c: f"{c['table'][6:]}-{c['N']}v{c['var']}")
@pytest.mark.parametrize("esz", [8, 4])
def test_pair_pass_slots_power_of_two(cfg, esz):
    """R2Pair::slot_k / slot_m == enc(rev(k)) / enc(rev(N - k)) for every thread and every unrolled m."""
    n, tpl = cfg["N"], cfg["TPL"]
    for t in range(tpl):
        e_up = enc(cfg, rev(cfg, t), esz)
        e_dn = enc(cfg, rev(cfg, tpl - t if t else 0), esz)
        for m in range((n // 2) // tpl):
            kc = m * tpl
            k = t + kc
            assert join(cfg, e_up, enc(cfg, rev(cfg, kc), esz)) == enc(cfg, rev(cfg, k), esz)
            c1 = enc(cfg, rev(cfg, n - tpl - kc), esz)
            c0 = enc(cfg, rev(cfg, (n - kc) % n), esz)
            km = (n - k) % n
            assert join(cfg, e_dn, c1 if t else c0) == enc(cfg, rev(cfg, km), esz)
            # global rows of the mirror: (TPL - t) + (N - TPL - kc)
            if k > 0:
                assert (tpl - t) + (n - tpl - kc) == n - k


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"{c['table'][6:]}-{c['N']}v{c['var']}")
def test_stage_slots_split_into_thread_and_unrolled_parts(cfg):
    """r2_dif_stage / r2_dit_stage: pos(t, m, q) = tpos(t) (+) cpos(m, q) wherever the kernels use the split."""
    n, tpl, rad = cfg["N"], cfg["TPL"], cfg["R"]
    pow2 = n & (n - 1) == 0
    e = n // tpl
    ns = n
    for rs in rad:
        L = ns // rs
        separable = pow2 or (cfg["ymode"] and (L % tpl == 0 if L >= tpl else tpl % L == 0))
        if separable:
            for t in range(tpl):
                tpos = t if L >= tpl else (t // L) * ns + (t % L)
                for m in range(e // rs):
                    u = t + tpl * m
                    base = (u // L) * ns + (u % L)
                    for q in range(rs):
                        cpos = ((tpl * m) // L) * ns + ((tpl * m) % L) + q * L
                        for esz in (8, 4):
                            assert join(cfg, enc(cfg, tpos, esz), enc(cfg, cpos, esz)) == enc(cfg, base + q * L, esz)
        ns = L
