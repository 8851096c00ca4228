#!/usr/bin/env python
"""bench.py -- Poisson-solve ns per grid point on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path, full grid, all host cores

One "step" = one `solver(...)` call (FFT x, FFT y, tridiagonal z, iFFT y, iFFT x) on the workload named in
`config.workload`: the channel grid of BASELINE.json's target (1024x512x512, periodic x/y, stretched Neumann-Neumann z,
FP64), which fits one GPU (2.15 GB per field).  Inputs are resident in HBM for `value`; `e2e` runs the same solve through the
public host-memory API (pinned host buffers, H2D + D2H inside the timed region).  The field (2.15 GB) is >> L2 (126 MB), so
no explicit L2 flush is needed between steps.  N > 1: ONE global grid as z slabs over the N GPUs (strong scaling); after the
timed region the distributed result is compared with a single-GPU solve of the same right-hand side (`parity`, exit != 0
on failure).

Other lines: `--impdiff` (one step = one RK substage of an is_impdiff run: Helmholtz solves of u, v, w + the Poisson solve;
`--dtdma-helmholtz` runs the three on the distributed TDMA), `--weak-nz K` (weak scaling: K planes per GPU), `--fp32`,
`--workload` (C2 / C4 / C5 ...), and the tuning switches of the library (`--dist-mode`, `--dist-windows`, `--chain-cols`,
`--pivot-dedup` ...).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ng, l, cbc, gr)
    "C3_channel_1024x512x512": ([1024, 512, 512], [12.0, 6.0, 2.0], [["P", "P"], ["P", "P"], ["N", "N"]], 2.0),
    "C2_tgv_512x512x512": ([512, 512, 512], [6.283185307179586] * 3, [["P", "P"]] * 3, 0.0),
    "C4_duct_1024x768x768": ([1024, 768, 768], [12.0, 2.0, 2.0], [["P", "P"], ["N", "N"], ["N", "N"]], 1.5),
    # configs[4]: 8 B200 (17.2 GB per field); the Poisson solve of the implicit-diffusion substep
    "C5_channel_2048x1024x1024": ([2048, 1024, 1024], [12.0, 6.0, 2.0], [["P", "P"], ["P", "P"], ["N", "N"]], 2.0),
    # diagnostic shape: at N = 2 every GPU sees the z pencil of C5 at N = 8 (2048 x 128 x 1024)
    "X_2048x256x1024": ([2048, 256, 1024], [12.0, 6.0, 2.0], [["P", "P"], ["P", "P"], ["N", "N"]], 2.0),
    "T_smoke_128x64x96": ([128, 64, 96], [6.0, 3.0, 2.0], [["P", "P"], ["P", "P"], ["N", "N"]], 2.0),   # harness smoke tests only
    "C1_ldc_2x64x64": ([2, 64, 64], [0.03125, 1.0, 1.0], [["P", "P"], ["N", "N"], ["N", "N"]], 0.0),
}
METRIC = "poisson_solve_ns_per_gridpoint"
UNIT = "ns/gridpoint"
ALGO_BYTES_PER_POINT_FP64 = 80.0   # 5 stages x (8 B read + 8 B write)
STAGE_BYTES_PER_POINT_FP64 = 16.0  # one stage: read + write the field once


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CpuArm:
    """The reference's CPU path for `config.workload` on the box's host cores: the oracle's threaded restatement
    (scipy/pocketfft r2r + the OpenMP C restatement of gaussel, oracle/gaussel_c.c), on the FULL grid.  The reference
    itself (Fortran + MPI + FFTW3) cannot be built in this image, so `kind` is "port".
    Imports nothing of the product package: only oracle/ is built and loaded."""

    def __init__(self, workload, threads=None):
        self.threads = int(threads or os.cpu_count() or 1)
        # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must use the cores it claims
        os.environ["OMP_NUM_THREADS"] = str(self.threads)
        import numpy as np
        from oracle.build import build_gaussel_c
        build_gaussel_c()
        from oracle import cans_oracle as O
        self.O, self.np = O, np
        self.workload = workload
        ng, l, cbc, gr = WORKLOADS[workload]
        self.ng, self.cbc = list(ng), cbc
        self.cs = O.make_case(ng, l, cbc, gr=gr)
        self.p = np.zeros((ng[2] + 2, ng[1] + 2, ng[0] + 2))
        self.rhs = O.hash_field(ng, 123)
        if all(b[0] in "PN" for b in cbc):   # compatible right-hand side, as the GPU arm's
            w = self.cs["dzf"][1:-1][:, None, None]
            self.rhs -= (self.rhs * w).sum() / (w.sum() * ng[0] * ng[1])
        self.npts = ng[0] * ng[1] * ng[2]

    def solve(self):
        """one `solver(...)` call on the full grid -> seconds"""
        cs, ng = self.cs, self.ng
        self.p[1:-1, 1:-1, 1:-1] = self.rhs
        t0 = time.perf_counter()
        self.O.solver_fast(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], self.cbc, ["c"] * 3,
                           self.p, workers=self.threads)
        return time.perf_counter() - t0


def cpu_baseline_leg(workload, seconds_target=20.0):
    """`cpu_baseline` of the product arm's line: a bounded sample (a few full-grid solves, ~10-30 s) on rank 0."""
    arm = CpuArm(workload)
    arm.solve()   # warm-up (page faults, pocketfft plan cache)
    times = []
    t0 = time.perf_counter()
    while len(times) < 2 or (time.perf_counter() - t0 < seconds_target and len(times) < 20):
        times.append(arm.solve())
    per = sorted(times)[len(times) // 2]
    return {"value": per * 1e9 / arm.npts, "unit": UNIT, "cores": arm.threads, "kind": "port",
            "sample": f"{len(times)} full-grid solves of {workload} ({'x'.join(map(str, arm.ng))}), median {per:.2f} s per solve, "
                      f"pocketfft workers = OMP threads = {arm.threads}"}


def bench_config(workload, world, extra=None):
    """`config` of both arms: the same keys and values, so that the driver's same_config check compares like with like"""
    ng, l, cbc, gr = WORKLOADS[workload]
    cfg = {"workload": workload, "grid": list(ng), "bc": "".join(b[0] + b[1] for b in cbc), "gr": gr,
           "l2": "field (>= 1 GB) larger than L2, no flush"}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args):
    """--impl reference: one step = one full-grid `solver(...)` call of the CPU path on all host cores (rank 0 only;
    under torchrun the other ranks exit 0 without work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args.workload)
    t_first = arm.solve()
    # keep the whole run within a few minutes: the warm-up solves are capped, the K timed steps always run
    warm_done = 1
    while warm_done < args.warmup and (warm_done + 1 + args.steps) * t_first < 240.0:
        arm.solve()
        warm_done += 1
    times = [arm.solve() for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    ns = ms * 1e6 / arm.npts
    out = {
        "impl": "reference", "metric": METRIC, "value": ns, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.workload, args.gpus),
        "note": "CPU restatement of src/solver.f90 (pocketfft r2r + OpenMP C gaussel) on the full grid; measured, not extrapolated; "
                f"warm-up solves run: {warm_done}",
        "cpu_baseline": {"value": ns, "unit": UNIT, "cores": arm.threads, "kind": "port",
                         "sample": f"{args.steps} full-grid solves of {args.workload}, mean {ms / 1e3:.2f} s per solve, "
                                   f"OMP_NUM_THREADS = pocketfft workers = {arm.threads}"},
        "e2e": {"value": ns, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def dist_parity(args, ctx, sd, cb, S, dist, torch, np, dev, rank, world, ng, dli, dzc, dzf, cbc, cf, fp32, tdt, ndt, step, p):
    """Driver-visible multi-GPU parity: one more distributed solve of a fresh right-hand side, the slabs gathered on
    rank 0 and compared with the solve of the SAME global right-hand side on rank 0's GPU alone, through the same C ABI
    (single-rank context).  Bar: 1e-12 relative L2 in FP64 (1e-5 in FP32); for a singular operator on a stretched grid
    the constant mode of the reference's own answer is rounding noise, so the metric that must pass is
    the one modulo the constant mode -- the strict number is printed beside it."""
    nl, I = ctx.n, (slice(1, -1),) * 3
    singular = all(b[0] in "PN" for b in cbc) and args.helmholtz == 0.0

    def make_rhs(c, t, n_loc, lo):
        S.fill_hash(c, t, n_loc, lo, 1, 4321)
        if all(b[0] in "PN" for b in cbc):
            z0 = lo[2] - 1
            wz = torch.from_numpy(dzf[1 + z0:1 + z0 + n_loc[2]].astype(ndt)).to(dev)[:, None, None]
            return (t[I] * wz).sum(), wz.sum() * ng[0] * ng[1]
        return None, None

    num, den = make_rhs(ctx, p, nl, ctx.lo)
    if num is not None:
        dist.all_reduce(num)
        dist.all_reduce(den)
        mean = float((num / den).item())
        p[I] -= mean
    step()
    torch.cuda.synchronize()
    assert ctx.dist_status() == 0, "a device-side wait timed out"
    # gather the slabs on rank 0 (whole planes incl. x / y halos: contiguous views)
    zs = [0]
    for r in range(world):
        zs.append(zs[-1] + ng[2] // world + (1 if r < ng[2] % world else 0))
    res = None
    if rank == 0:
        full = torch.zeros((ng[2] + 2, ng[1] + 2, ng[0] + 2), dtype=tdt, device=dev)
        full[1:1 + nl[2]] = p[1:-1]
        for r in range(1, world):
            dist.recv(full[1 + zs[r]:1 + zs[r + 1]], src=r)
        ctx1 = cb.Context(ng, is_fp32=fp32)
        sd1 = cb.initsolver(ctx1, ng, dli, 1.0 / dzc, 1.0 / dzf, cbc, [[0.0, 0.0]] * 3, cf, device=dev)
        q = torch.empty_like(full)
        n1, d1 = make_rhs(ctx1, q, ng, [1, 1, 1])
        if n1 is not None:
            q[I] -= mean   # the very same scalar the slabs subtracted
        if args.helmholtz != 0.0:
            alphai = 1.0 / args.helmholtz
            cb.solver(ng, ng, sd1.arrplan, float(sd1.normfft) * alphai, sd1.lambdaxy, sd1.a, sd1.b + alphai, sd1.c, cbc, cf, q)
        else:
            cb.solver(ng, ng, sd1.arrplan, sd1.normfft, sd1.lambdaxy, sd1.a, sd1.b, sd1.c, cbc, cf, q)
        torch.cuda.synchronize()
        a, b = full[I].double(), q[I].double()
        strict = float(((a - b).norm() / b.norm()).item())
        dm = (a - b) - (a - b).mean()
        modc = float((dm.norm() / (b - b.mean()).norm()).item())
        tol = 1e-5 if fp32 else 1e-12
        err = modc if singular else strict
        res = {"vs": "single-GPU solve of the same global right-hand side through the same C ABI (rank 0)",
               "rel_l2": err, "rel_l2_strict": strict, "rel_l2_mod_const": modc, "tol": tol,
               "metric": "modulo the constant mode (singular operator)" if singular else "strict", "ok": bool(err < tol),
               "dtdma": bool(args.dtdma)}
        if args.dtdma:
            # the distributed TDMA is a different (equally valid) elimination order: the reference's own gaussel_dtdma differs from
            # its gaussel by the conditioning of the system, not by 1e-12 (tests compare it with the DTDMA oracle instead)
            res["tol"] = tol = 1e-9 if not fp32 else 1e-4
            res["ok"] = bool(err < tol)
        sd1.arrplan.destroy()
        ctx1.close()
        del full, q
    else:
        dist.send(p[1:-1].contiguous(), dst=0)
    box = [res]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cans_b200", choices=["cans_b200", "reference"])
    ap.add_argument("--workload", default="C3_channel_1024x512x512", choices=sorted(WORKLOADS))
    ap.add_argument("--thomas", type=int, default=-1)
    ap.add_argument("--fft-x-lines", type=int, default=-1)
    ap.add_argument("--fft-y-lines", type=int, default=-1)
    ap.add_argument("--r2-flags", type=int, default=-1)
    ap.add_argument("--x-variant", type=int, default=-1)
    ap.add_argument("--y-variant", type=int, default=-1)
    ap.add_argument("--chain-cols", type=int, default=-1)
    ap.add_argument("--chain-streams", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fp32", action="store_true", help="the -D_SINGLE_PRECISION build of CaNS (not the headline; FP64 is)")
    ap.add_argument("--host-chunks", type=int, default=16)
    ap.add_argument("--zmajor", type=int, default=-1)
    ap.add_argument("--pivot-dedup", type=int, default=-1, help="0 = full pivot cache, -1 = library default (deduplicated where lambda is symmetric)")
    ap.add_argument("--dtdma", action="store_true",
                    help="N > 1: the distributed-TDMA path (CANSB200_CTX_DTDMA); use it with --helmholtz (it has no singular-pivot pin)")
    ap.add_argument("--dist-windows", type=int, default=-1, help="N > 1: x windows of the pipelined exchange (-1 auto, 1 = two barriers)")
    ap.add_argument("--dist-thomas-ctas", type=int, default=-1, help="N > 1: CTAs of the tridiagonal kernel inside the pipeline (-1 auto)")
    ap.add_argument("--dist-mode", type=int, default=-1, help="N > 1: 0 = peer stores of the producing kernels, 1 = copy engines, -1 = auto")
    ap.add_argument("--dist-chunks", type=int, default=-1, help="N > 1, copy engines: z chunks of the forward half (-1 auto)")
    ap.add_argument("--dist-split-pad", type=int, default=-1, help="N > 1, mode 2: KB of shared-memory padding of the forward y transforms")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the check of the distributed result against a single-GPU solve")
    ap.add_argument("--impdiff", action="store_true",
                    help="one step = one RK substage of an is_impdiff run (src/main.f90:447-467): the Helmholtz solves of u, v, w "
                         "(face-centred in x / y / z, three alternating alpha) + the pressure Poisson solve; value = ns per point per SOLVE")
    ap.add_argument("--dtdma-helmholtz", action="store_true",
                    help="--impdiff, N > 1: the three Helmholtz solves run on a distributed-TDMA context (is_poisson_dtdma), the Poisson "
                         "solve (singular operator) stays on the transposes")
    ap.add_argument("--weak-nz", type=int, default=0, help="weak scaling: nz = this many planes per GPU (the workload's nx, ny are kept)")
    ap.add_argument("--helmholtz", type=float, default=0.0,
                    help="alpha != 0: time the implicit-diffusion Helmholtz solve p/alpha + lap p = rhs (regular operator) instead")
    ap.add_argument("--fillps", action="store_true",
                    help="one step = the pressure-correction right-hand side + its solve (src/main.f90:465-467: fillps, updt_rhs_b, solver) "
                         "as ONE call, the forward x transform evaluating fillps at load time (cansb200_solve_fillps); the split sequence "
                         "(fillps kernel, then solver) is timed beside it.  One GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import cans_b200 as cb
    S = sys.modules["cans_b200.solver"]

    ng, l, cbc, gr = WORKLOADS[args.workload]
    if args.weak_nz > 0:
        ng = [ng[0], ng[1], args.weak_nz * world]
    cf = ["c"] * 3
    # grid + operator exactly as initgrid/initsolver build them (host arithmetic, as in the reference)
    per_z = cbc[2] == ["P", "P"]
    from cans_b200 import gridgen
    dzc, dzf = gridgen.initgrid(1, ng[2], gr, l[2], per_z)
    dli = [ng[0] / l[0], ng[1] / l[1], ng[2] / l[2]]
    # N > 1: ONE global grid, z slabs over the N GPUs of the box (strong scaling, BASELINE.json "1/2/4/8 B200")
    fp32 = bool(args.fp32)
    tdt, ndt, esz = (torch.float32, np.float32, 4) if fp32 else (torch.float64, np.float64, 8)
    ctx = cb.Context(ng, is_fp32=fp32, rank=rank, nranks=world)
    ctx.connect()
    if args.dtdma and world > 1:
        ctx.set_dtdma(True)
    ctx.set_dist_windows(args.dist_windows, args.dist_thomas_ctas)
    ctx.set_dist_mode(args.dist_mode, args.dist_chunks, args.dist_split_pad)
    ctx.set_variant(args.x_variant, args.y_variant)
    if args.r2_flags >= 0:
        ctx.set_r2_flags(args.r2_flags)
    ctx.set_chain(args.chain_cols, args.chain_streams)
    if args.zmajor >= 0:
        ctx.set_zmajor(bool(args.zmajor))
    sd = cb.initsolver(ctx, ng, dli, 1.0 / dzc, 1.0 / dzf, cbc, [[0.0, 0.0]] * 3, cf, device=dev,
                       thomas_variant=args.thomas, fft_x_lines=args.fft_x_lines, fft_y_lines=args.fft_y_lines,
                       pivot_dedup=args.pivot_dedup)
    nl = ctx.n                      # local x pencil (nx, ny, nz / N)
    shp = (nl[2] + 2, nl[1] + 2, nl[0] + 2)
    npts = ng[0] * ng[1] * ng[2]    # global points
    npts_local = nl[0] * nl[1] * nl[2]
    p = torch.empty(shp, dtype=tdt, device=dev)
    S.fill_hash(ctx, p, nl, ctx.lo, 1, 123)   # hash of the GLOBAL index: every decomposition sees the same field
    I = (slice(1, -1),) * 3
    if all(b[0] in "PN" for b in cbc):
        z0 = ctx.lo[2] - 1
        wz = torch.from_numpy(dzf[1 + z0:1 + z0 + nl[2]].astype(ndt)).to(dev)[:, None, None]
        num = (p[I] * wz).sum()
        den = wz.sum() * ng[0] * ng[1]
        if world > 1:
            dist.all_reduce(num)
            dist.all_reduce(den)
        p[I] -= num / den

    nsolves_per_step = 1
    if args.impdiff:
        # the velocity operators of the same flow: a pressure Neumann wall is a velocity Dirichlet wall; u / v / w are face
        # centred in x / y / z (src/main.f90:296-316)
        cbcv = [["D", "D"] if b == ["N", "N"] else list(b) for b in cbc]
        ctx_h = ctx
        if args.dtdma_helmholtz and world > 1:
            ctx_h = cb.Context(ng, is_fp32=fp32, rank=rank, nranks=world)
            ctx_h.connect()
            ctx_h.set_dtdma(True)
        vel = []
        for d in range(3):
            cfv = ["c"] * 3
            cfv[d] = "f"
            sdv = cb.initsolver(ctx_h, ng, dli, 1.0 / dzc, 1.0 / dzf, cbcv, [[0.0, 0.0]] * 3, cfv, device=dev, cache_slots=3)
            f = torch.empty(shp, dtype=tdt, device=dev)
            S.fill_hash(ctx, f, nl, ctx.lo, 1, 1000 + d)
            vel.append((sdv, cfv, f, f.clone()))
        p_init = p.clone()
        visc_dt = 1.0e-4
        alphas = [-0.5 * visc_dt * r for r in (32.0 / 60.0, 8.0 / 60.0, 20.0 / 60.0)]   # rkcoeff sums of the three substeps
        nsolves_per_step = 4
        irk = [0]

        def step():
            al = alphas[irk[0] % 3]
            irk[0] += 1
            for sdv, cfv, f, _ in vel:
                cb.solve_helmholtz(ctx_h.n, ng, ctx_h.hi(), sdv.arrplan, sdv.normfft, al, sdv.lambdaxy, sdv.a, sdv.b, sdv.c, None, None,
                                   None, ctx_h.is_bound(), cbcv, cfv, f)
            cb.solver(nl, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, p)
    elif args.fillps:
        if world > 1:
            raise SystemExit("bench.py --fillps: one GPU (the z halos of u, v, w would need the host's halo exchange)")
        # a prediction velocity with consistent halos: periodic directions wrap, no flow through Neumann (pressure) walls
        uvw = []
        for d in range(3):
            f = torch.empty(shp, dtype=tdt, device=dev)
            S.fill_hash(ctx, f, nl, ctx.lo, 1, 2000 + d)
            uvw.append(f)
        for d, ax in ((0, 2), (1, 1), (2, 0)):      # direction d of the grid is tensor axis ax
            for f in uvw:
                lo_h, hi_h = [slice(None)] * 3, [slice(None)] * 3
                lo_i, hi_i = [slice(None)] * 3, [slice(None)] * 3
                lo_h[ax], hi_h[ax], lo_i[ax], hi_i[ax] = 0, -1, 1, -2
                if cbc[d] == ["P", "P"]:
                    f[tuple(lo_h)] = f[tuple(hi_i)]
                    f[tuple(hi_h)] = f[tuple(lo_i)]
            if cbc[d] != ["P", "P"]:
                f = uvw[d]                           # the wall-normal component vanishes on both walls
                lo_h, hi_i = [slice(None)] * 3, [slice(None)] * 3
                lo_h[ax], hi_i[ax] = 0, -2
                f[tuple(lo_h)] = 0.0
                f[tuple(hi_i)] = 0.0
        dzfi_d = torch.from_numpy((1.0 / dzf).astype(ndt)).to(dev)
        dti_f = 1.0e3

        def step():
            S.solver_fillps(nl, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, dli, dzfi_d, dti_f,
                            uvw[0], uvw[1], uvw[2], p)
    elif args.helmholtz != 0.0:
        alphai = 1.0 / args.helmholtz
        b_h, norm_h = sd.b + alphai, float(sd.normfft) * alphai   # src/solve_helmholtz.f90:63-71

        def step():
            cb.solver(nl, ng, sd.arrplan, norm_h, sd.lambdaxy, sd.a, b_h, sd.c, cbc, cf, p)
    else:
        def step():
            cb.solver(nl, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, p)

    # FP32 only: the null-space mode of the singular operator is not pinned on a stretched grid, so
    # feeding a solution back as the next right-hand side overflows single precision after a few solves.  The FP32
    # line therefore restores the right-hand side before every solve (outside the per-solve CUDA events).
    p0 = p.clone() if fp32 else None
    per_step_events = fp32 or args.impdiff   # the fields are restored before every step, outside the per-step events

    def reset():
        if fp32:
            p.copy_(p0)
        if args.impdiff:   # a Helmholtz solve with |alpha| ~ 1e-5 shrinks its field by that factor: restore the right-hand sides
            p.copy_(p_init)
            for _, _, f, f0 in vel:
                f.copy_(f0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        reset()
        step()
    sync_all()
    l0 = sd.arrplan.stats()["launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    if not per_step_events:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        sync_all()
        ms_total = ev0.elapsed_time(ev1)
    else:
        evs = []
        for _ in range(args.steps):
            reset()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            step()
            a1.record()
            evs.append((a0, a1))
        sync_all()
        ms_total = sum(a0.elapsed_time(a1) for a0, a1 in evs)
    clocks = sampler.stop() if rank == 0 else None
    launches = sd.arrplan.stats()["launches"] - l0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    assert bool(torch.isfinite(p[I]).all()), "solution is not finite"
    assert ctx.dist_status() == 0, "a device-side barrier timed out"
    total_pts = npts  # one global grid, whatever N
    value = ms_per_step * 1e6 / total_pts / nsolves_per_step

    # ---- per-stage device times (live, CUDA events on the solve's stream) -> roofline of the dominant kernel
    ctx.set_profiling(True)
    for _ in range(args.steps):
        reset()
        step()
    prof, nprof = ctx.get_profile()
    ctx.set_profiling(False)
    stage_ms = {k: v / max(nprof, 1) for k, v in prof.items()}
    peak, peak_src = measured_peaks()
    heavy = {k: v for k, v in stage_ms.items() if k != "pivot_cache"}
    dom = max(heavy, key=heavy.get)
    if world > 1:   # max over ranks of every stage time
        keys = sorted(stage_ms)
        t = torch.tensor([stage_ms[k] for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage_ms = {k: float(v) for k, v in zip(keys, t.tolist())}
        heavy = {k: v for k, v in stage_ms.items() if k != "pivot_cache"}
        dom = max(heavy, key=heavy.get)
    stage_bytes = STAGE_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local   # per GPU: each rank's launch covers its local points
    fused_x = args.fillps and dom == "fft_x_fwd"
    if fused_x:   # the fused forward x transform reads u, v, w (24 B/point) and writes the field (8)
        stage_bytes = 32.0 * (esz / 8.0) * npts_local
    achieved = stage_bytes / (stage_ms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = None if fp32 else json.load(open(tpath)).get(args.workload, {}).get("fft_x_fwd_fillps" if fused_x else dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": stage_bytes,
                "stage_ms": stage_ms,
                "solve": {"algorithmic_bytes_per_gpu": ALGO_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local,
                          "achieved": nsolves_per_step * ALGO_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local / (ms_per_step * 1e-3) / 1e9,
                          "frac": nsolves_per_step * ALGO_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local / (ms_per_step * 1e-3) / 1e9 / peak}}
    if world > 1:
        # two exchanges per solve, each sends (P-1)/P of the local field per GPU per direction
        nvb = 2.0 * (world - 1) / world * npts_local * float(esz)
        ach = nsolves_per_step * nvb / (ms_per_step * 1e-3) / 1e9   # (a distributed-TDMA solve moves far less: upper bound then)
        roofline["nvlink"] = {"bytes_per_gpu_per_direction_per_solve": nvb, "achieved_GBs_over_whole_solve": ach,
                              "peak_GBs": 900.0, "frac": ach / 900.0, "measured_peer_copy_GBs": 770.0, "frac_of_measured": ach / 770.0,
                              "floor_ms_at_900": nvb / 900e9 * 1e3,
                              "note": "peer stores are issued by the y-transform and tridiagonal kernels themselves; achieved = NVLink "
                                      "bytes / WHOLE solve time (the exchange overlaps the local stages window by window); SM stores "
                                      "reach 690 GB/s in any pattern on this box (profiles/r2a_nvlink_store_patterns.txt)"}
    roofline["stage_ms_mode"] = ("per-stage times come from a profiling pass in which the stages run back to back on one stream "
                                 "(no x windows / second stream, N > 1: one window with two whole-field barriers); `value` is the "
                                 "default overlapped schedule")

    # ---- --fillps: the same step as three separate launches sequences (fillps kernel, then the solve reading p back)
    fillps_info = None
    if args.fillps:
        fused_flag = int(sd.arrplan.stats()["fillps_fused"])
        ctx.set_fuse_fillps(False)
        for _ in range(3):
            step()
        sync_all()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            step()
        b1.record()
        sync_all()
        ctx.set_fuse_fillps(True)
        split_ms = b0.elapsed_time(b1) / args.steps
        fillps_info = {"fused": fused_flag, "fused_ms_per_step": ms_per_step, "split_ms_per_step": split_ms,
                       "algorithmic_bytes_per_point": {"fused": 96.0 * esz / 8.0, "split": 112.0 * esz / 8.0},
                       "note": "step = fillps + solver of the pressure correction (src/main.f90:465-467); fused: the forward x "
                               "transform reads u, v, w (24 B/pt) instead of p (8), and the fillps kernel (24 read + 8 written) is gone"}

    # ---- e2e: the host-memory API (mode A of SURVEY 8b): pinned host p, H2D + solve + D2H per step
    e2e = None
    if not args.no_e2e and not args.impdiff and not args.fillps:
        nb = int(np.prod(shp)) * esz
        if world == 1:
            ctx.set_host_chunks(args.host_chunks)
            if args.host_chunks > 1:   # chunked pipeline: only the interior z planes travel (halo planes are never read)
                nb = nl[2] * shp[1] * shp[2] * esz
        ph = torch.empty(shp, dtype=tdt).pin_memory()
        ph.copy_(p)
        pn = ph.numpy()
        hs = sd.host
        ne2e = max(1, min(args.steps, 5))
        cb.solver(nl, ng, sd.arrplan, sd.normfft, hs["lambdaxy"], hs["a"], hs["b"], hs["c"], cbc, cf, pn)  # warm-up
        sync_all()
        el = 0.0
        ph0 = ph.clone() if fp32 else None
        for _ in range(ne2e):
            if fp32:
                ph.copy_(ph0)
            t0 = time.perf_counter()
            cb.solver(nl, ng, sd.arrplan, sd.normfft, hs["lambdaxy"], hs["a"], hs["b"], hs["c"], cbc, cf, pn)
            torch.cuda.synchronize()
            el += time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": el / ne2e * 1e9 / total_pts, "unit": UNIT,
               "h2d_bytes_per_step": (nb + esz * (3 * ng[2] + ctx.n_z[0] * ctx.n_z[1])) * world,
               "d2h_bytes_per_step": nb * world, "steps": ne2e, "ms_per_step": el / ne2e * 1e3,
               "host_chunks": args.host_chunks if world == 1 else 1}
        del ph

    # ---- N > 1: the distributed result against a single-GPU solve of the same global right-hand side (rank 0)
    parity = None
    if world > 1 and not args.no_parity and not args.impdiff and args.weak_nz == 0:
        parity = dist_parity(args, ctx, sd, cb, S, dist, torch, np, dev, rank, world, ng, dli, dzc, dzf, cbc, cf, fp32, tdt, ndt, step, p)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not fp32 and not args.fillps:
        cpu = cpu_baseline_leg(args.workload)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak" if args.weak_nz > 0 else "strong",
            "vs_baseline": None, "dtype": "f32" if fp32 else "f64", "data": "synthetic",
            "config": bench_config(args.workload, world),   # identical in both arms (driver's same_config)
            "details": {
                "decomposition": "single GPU" if world == 1 else f"z slabs over {world} GPUs (x pencils, dims=[1,{world}]), "
                                 "transposes = peer-mapped stores over NVLink, pipelined over x windows",
                "solves_per_s": 1e3 / ms_per_step, "thomas_variant": int(sd.arrplan.stats()["thomas_variant"]),
                **({"dtdma": True} if (args.dtdma and world > 1) else {}),
                **({"impdiff": "one step = Helmholtz solves of u, v, w (three alternating alpha) + the Poisson solve; value is per solve",
                    "solves_per_step": 4, "helmholtz_path": "distributed TDMA" if (args.dtdma_helmholtz and world > 1) else "transposes"}
                   if args.impdiff else {}),
                **({"weak_nz_per_gpu": args.weak_nz, "grid_solved": ng} if args.weak_nz > 0 else {}),
                **({"fillps": fillps_info} if fillps_info is not None else {}),
                **({"helmholtz_alpha": args.helmholtz} if args.helmholtz != 0.0 else {})},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if parity is not None:
            out["parity"] = parity
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"bench.py: distributed result differs from the single-GPU solve: {parity}")


if __name__ == "__main__":
    main()
