#!/usr/bin/env python
"""bench.py -- Poisson-solve ns per grid point on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

One "step" = one `solver(...)` call (FFT x, FFT y, tridiagonal z, iFFT y, iFFT x)
on the workload named in `config.workload`: the channel grid of BASELINE.json's
target (1024x512x512, periodic x/y, stretched Neumann-Neumann z, FP64), which
fits one GPU (2.15 GB per field).  Inputs are resident in HBM for `value`; `e2e`
runs the same solve through the public host-memory API (pinned host buffers,
H2D + D2H inside the timed region).  The field (2.15 GB) is >> L2 (126 MB), so
no explicit L2 flush is needed between steps.

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


def cpu_solve_sample(workload, seconds_target=15.0, threads=None):
    """The oracle's threaded restatement (scipy/pocketfft r2r + OpenMP C gaussel) on a bounded
    x-y sub-sample of the workload: same nz, same BCs, same stretched grid, nx*ny reduced."""
    import numpy as np
    from oracle import cans_oracle as O
    import __graft_entry__ as g
    g.build()
    ng_full, l, cbc, gr = WORKLOADS[workload]
    threads = threads or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    # sample: keep nz, shrink nx and ny by the same power of two until ~64 Mi points
    ng = list(ng_full)
    while ng[0] * ng[1] * ng[2] > 2 ** 25 and ng[0] > 64 and ng[1] > 32:
        ng[0] //= 2
        ng[1] //= 2
    ls = [l[0] * ng[0] / ng_full[0], l[1] * ng[1] / ng_full[1], l[2]]
    cs = O.make_case(ng, ls, cbc, gr=gr)
    p = np.zeros((ng[2] + 2, ng[1] + 2, ng[0] + 2))
    p[1:-1, 1:-1, 1:-1] = O.hash_field(ng, 123)

    def one():
        O.solver_fast(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cbc, ["c"] * 3, p,
                      workers=threads)
    one()  # warm-up (plans, page faults)
    t0 = time.perf_counter()
    reps = 0
    while True:
        one()
        reps += 1
        el = time.perf_counter() - t0
        if el > seconds_target or reps >= 50:
            break
    per = el / reps
    npts = ng[0] * ng[1] * ng[2]
    return {"ns_per_point": per * 1e9 / npts, "sample": f"{ng[0]}x{ng[1]}x{ng[2]} sub-grid of {workload} "
            f"(same nz, BCs, stretching), {reps} solves in {el:.1f} s", "cores": threads, "sec_per_solve": per, "ng": ng}


def run_reference(args):
    """--impl reference: the reference's CPU path (restated: the reference needs gfortran+MPI+FFTW3,
    none of which exist in this image) on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    total = max(1, args.warmup + args.steps)
    # bound the whole run to a few minutes
    per_step_budget = min(10.0, 150.0 / total)
    res = []
    for _ in range(args.warmup):
        cpu_solve_sample(args.workload, seconds_target=0.0, threads=threads)
    for _ in range(args.steps):
        res.append(cpu_solve_sample(args.workload, seconds_target=per_step_budget * 0.5, threads=threads))
    ns = sorted(r["ns_per_point"] for r in res)[len(res) // 2]
    ng_full = WORKLOADS[args.workload][0]
    npts = ng_full[0] * ng_full[1] * ng_full[2]
    out = {
        "impl": "reference", "metric": METRIC, "value": ns, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ns * npts * 1e-6, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "grid": ng_full, "note": "CPU restatement of solver.f90 (pocketfft r2r + "
                   "OpenMP C gaussel); ms_per_step extrapolated from the sampled ns/gridpoint to the full grid"},
        "cpu_baseline": {"value": ns, "unit": UNIT, "cores": threads, "kind": "port", "sample": res[0]["sample"]},
        "e2e": {"value": ns, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


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
    ap.add_argument("--dtdma", action="store_true",
                    help="N > 1: the distributed-TDMA path (CANSB200_CTX_DTDMA); use it with --helmholtz (it has no singular-pivot pin)")
    ap.add_argument("--helmholtz", type=float, default=0.0,
                    help="alpha != 0: time the implicit-diffusion Helmholtz solve p/alpha + lap p = rhs (regular operator) instead")
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
    ctx.set_variant(args.x_variant, args.y_variant)
    if args.r2_flags >= 0:
        ctx.set_r2_flags(args.r2_flags)
    ctx.set_chain(args.chain_cols, args.chain_streams)
    if args.zmajor >= 0:
        ctx.set_zmajor(bool(args.zmajor))
    sd = cb.initsolver(ctx, ng, dli, 1.0 / dzc, 1.0 / dzf, cbc, [[0.0, 0.0]] * 3, cf, device=dev,
                       thomas_variant=args.thomas, fft_x_lines=args.fft_x_lines, fft_y_lines=args.fft_y_lines)
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

    if args.helmholtz != 0.0:
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

    def reset():
        if fp32:
            p.copy_(p0)

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
    if not fp32:
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
    value = ms_per_step * 1e6 / total_pts

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
    achieved = stage_bytes / (stage_ms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = None if fp32 else json.load(open(tpath)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": stage_bytes,
                "stage_ms": stage_ms,
                "solve": {"algorithmic_bytes_per_gpu": ALGO_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local,
                          "achieved": ALGO_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local / (ms_per_step * 1e-3) / 1e9,
                          "frac": ALGO_BYTES_PER_POINT_FP64 * (esz / 8.0) * npts_local / (ms_per_step * 1e-3) / 1e9 / peak}}
    if world > 1:
        # two exchanges per solve, each sends (P-1)/P of the local field per GPU per direction
        nvb = 2.0 * (world - 1) / world * npts_local * float(esz)
        roofline["nvlink"] = {"bytes_per_gpu_per_direction_per_solve": nvb, "peak_GBs": 770.0,
                              "peak_source": "B200_PROFILING.md measured peer copy (900 nominal)",
                              "floor_ms": nvb / 770e9 * 1e3,
                              "note": "peer stores are issued by the y-transform and tridiagonal kernels themselves; "
                                      "their stage times above include the wire time"}

    # ---- e2e: the host-memory API (mode A of SURVEY 8b): pinned host p, H2D + solve + D2H per step
    e2e = None
    if not args.no_e2e:
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not fp32:
        r = cpu_solve_sample(args.workload, seconds_target=15.0)
        cpu = {"value": r["ns_per_point"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if fp32 else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "grid": ng, "bc": "".join(b[0] + b[1] for b in cbc), "gr": gr,
                       "decomposition": "single GPU" if world == 1 else f"z slabs over {world} GPUs (x pencils, dims=[1,{world}]), "
                                        "transposes = peer-mapped stores over NVLink",
                       "l2": "inputs (2.15 GB/field) larger than L2, no flush", "solves_per_s": 1e3 / ms_per_step,
                       "thomas_variant": int(sd.arrplan.stats()["thomas_variant"]),
                       **({"dtdma": True} if (args.dtdma and world > 1) else {}),
                       **({"helmholtz_alpha": args.helmholtz} if args.helmholtz != 0.0 else {})},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
