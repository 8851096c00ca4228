#!/usr/bin/env python
"""Per-stage device timings of the solver kernels through the stage-level C ABI
(cansb200_r2r / cansb200_gaussel), for every tuning variant of the fast transforms.

    python scripts/bench_stages.py [--grid 1024 512 512] [--reps 5]

Prints ms, GB/s (16 B per point per stage, FP64) and the fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cans_b200 as cb  # noqa: E402
from cans_b200 import gridgen  # noqa: E402

S = sys.modules["cans_b200.solver"]
KINDS = {"R2HC": 0, "HC2R": 1, "REDFT01": 4, "REDFT10": 5, "RODFT01": 8, "RODFT10": 9}


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[1024, 512, 512])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kinds", default="R2HC,HC2R")
    ap.add_argument("--variants", type=int, default=3)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--no-thomas", action="store_true")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    ng = args.grid
    dev = torch.device("cuda:0")
    peak = 6542.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dt = torch.float32 if args.fp32 else torch.float64
    esz = 4 if args.fp32 else 8
    ctx = cb.Context(ng, is_fp32=args.fp32)
    ctx.set_r2_flags(args.flags)
    arr = torch.rand((ng[2], ng[1], ng[0]), dtype=dt, device=dev)
    npts = ng[0] * ng[1] * ng[2]
    rows = []
    for axis, n in ((0, ng[0]), (1, ng[1])):
        for kname in args.kinds.split(","):
            for var in range(args.variants):
                ctx.set_variant(var if axis == 0 else 0, var if axis == 1 else 0)
                if args.check:
                    # every variant must reproduce variant 0 on the same input (a small slab is enough)
                    small = [ng[0], ng[1], 4]
                    cs = cb.Context(small, is_fp32=args.fp32)
                    x0 = torch.rand((4, ng[1], ng[0]), dtype=dt, device=dev) - 0.5
                    a0, a1 = x0.clone(), x0.clone()
                    cs.set_variant(0, 0)
                    S.r2r(cs, KINDS[kname], n, axis, a0)
                    cs.set_variant(var if axis == 0 else 0, var if axis == 1 else 0)
                    S.r2r(cs, KINDS[kname], n, axis, a1)
                    err = float((a1 - a0).abs().max() / a0.abs().max())
                    assert err < (1e-5 if args.fp32 else 1e-13), (kname, axis, var, err)
                ms = timeit(lambda: S.r2r(ctx, KINDS[kname], n, axis, arr), args.reps)
                rows.append((f"r2r axis={axis} n={n} {kname} var={var}", ms, 2 * esz * npts / ms / 1e6))
    ctx.set_variant(0, 0)
    # Thomas (Neumann-Neumann stretched, and periodic)
    for cbcz, per in (() if args.no_thomas else ((["N", "N"], False), (["P", "P"], True))):
        cbc = [["P", "P"], ["P", "P"], cbcz]
        dzc, dzf = gridgen.initgrid(1, ng[2], 0.0 if per else 2.0, 2.0, per)
        dli = [ng[0] / 12.0, ng[1] / 6.0, ng[2] / 2.0]
        sd = cb.initsolver(ctx, ng, dli, 1.0 / dzc, 1.0 / dzf, cbc, [[0.0, 0.0]] * 3, ["c"] * 3, device=dev)
        ms = timeit(lambda: S.gaussel(sd.arrplan, ng[2], sd.a, sd.b, sd.c, per, 1.0, arr, sd.lambdaxy), args.reps)
        rows.append((f"gaussel nz={ng[2]} periodic={per}", ms, 2 * esz * npts / ms / 1e6))
        sd.arrplan.destroy()
    for name, ms, gbs in rows:
        print(f"{name:44s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.3f} of measured HBM peak")


if __name__ == "__main__":
    main()
