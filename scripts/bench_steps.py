#!/usr/bin/env python
"""Device times of the steps either side of the solve (fillps, correc, chkdiv) on the C3 grid, both launch geometries
(CANSB200_CTX_AUX_3D).  Prints one JSON line.  Run on the GPU box: python scripts/bench_steps.py [nx ny nz]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cans_b200 as cb
from cans_b200 import gridgen

S = sys.modules["cans_b200.solver"]
ng = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [1024, 512, 512]
l = [12.0, 6.0, 2.0]
dev = torch.device("cuda:0")
dzc, dzf = gridgen.initgrid(1, ng[2], 2.0, l[2], False)
dli = [ng[d] / l[d] for d in range(3)]
ctx = cb.Context(ng)
shp = (ng[2] + 2, ng[1] + 2, ng[0] + 2)
u, v, w, p = (torch.empty(shp, dtype=torch.float64, device=dev) for _ in range(4))
for t, seed in ((u, 1), (v, 2), (w, 3), (p, 4)):
    S.fill_hash(ctx, t, ng, [1, 1, 1], 1, seed)
dzfi, dzci = torch.from_numpy(1.0 / dzf).to(dev), torch.from_numpy(1.0 / dzc).to(dev)
npts = ng[0] * ng[1] * ng[2]
out = {"grid": ng, "reps": 10, "bytes_per_point": {"fillps": 32, "correc": 56, "chkdiv": 24}}
for geom in (0, 1):
    ctx.set_aux_3d(bool(geom))
    res = {}
    for name, fn in (("fillps", lambda: S.fillps(ctx, ng, dli, dzfi, 1.0e3, u, v, w, p)),
                     ("correc", lambda: S.correc(ctx, ng, dli, dzci, 1.0e-9, p, u, v, w))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = {"ms": ms, "GBs": out["bytes_per_point"][name] * npts / ms / 1e6}
    S.chkdiv(ctx, ng, l, dli, dzfi, u, v, w)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        S.chkdiv(ctx, ng, l, dli, dzfi, u, v, w)   # synchronous: returns the two reductions
    ms = (time.perf_counter() - t0) / 5 * 1e3
    res["chkdiv"] = {"ms_wall": ms, "GBs": out["bytes_per_point"]["chkdiv"] * npts / ms / 1e6}
    out["geometry_3d" if geom else "flat_index"] = res
print(json.dumps(out))
