"""Distributed (z-slab) solve against the single-rank oracle.  Launched by tests/test_gpu_dist.py as

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_check.py

Every rank solves its slab of the same global problem through the C ABI (peer-mapped exchange over
NVLink) and compares it with the oracle's global solution.  Prints `DIST_OK <max err>` on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import cans_b200 as cb  # noqa: E402
from cans_b200.decomp import SlabDecomp  # noqa: E402

DIST_CASES = cases.DIST_CASES
P = cases.P


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")
    from oracle import cans_oracle as O
    worst = 0.0
    ndtdma = 0
    for name, (ng, l, cbc, cf, gr, dt, helm) in DIST_CASES.items():
        cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
        p = cases.make_rhs(cs)
        ref = p.copy()
        if helm:
            O.solve_helmholtz(ng, ng, cs["arrplan"], cs["normfft"], cases.ALPHA, cs["lambdaxy"], cs["a"], cs["b"], cs["c"],
                              None, None, None, cbc, cf, ref)
        else:
            O.solver(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cbc, cf, ref)
        ctx = cb.Context(ng, is_fp32=dt == np.float32, rank=rank, nranks=world)
        dec = SlabDecomp(ng, world, rank)
        assert ctx.n == dec.n and ctx.lo == dec.lo and ctx.n_z == dec.n_z and ctx.lo_z == dec.lo_z, (ctx.n, dec.n)
        ctx.connect()
        sd = cb.initsolver(ctx, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev)
        z0, z1 = dec.z_range()
        pl = np.zeros((z1 - z0 + 2, ng[1] + 2, ng[0] + 2), dtype=dt)
        pl[1:-1] = p[1 + z0:1 + z1]
        pd = torch.from_numpy(pl).to(dev)
        for rep in range(3):   # repeated solves exercise buffer reuse across the device-side barriers
            pd.copy_(torch.from_numpy(pl))
            if helm:
                cb.solve_helmholtz(ctx.n, ng, ctx.hi(), sd.arrplan, sd.normfft, cases.ALPHA, sd.lambdaxy, sd.a, sd.b, sd.c, None, None,
                                   None, ctx.is_bound(), cbc, cf, pd)
            else:
                cb.solver(ctx.n, ng, sd.arrplan, sd.normfft, sd.lambdaxy, sd.a, sd.b, sd.c, cbc, cf, pd)
        torch.cuda.synchronize()
        assert ctx.dist_status() == 0, "a device-side barrier timed out"
        got = pd.cpu().numpy()
        # parity metric on the GLOBAL field: gather the slabs (the null-space rule needs the global mean)
        parts = [None] * world
        dist.all_gather_object(parts, got[1:-1, 1:-1, 1:-1])
        full = np.concatenate(parts, axis=0)
        err = cases.parity_error(cs, full, ref[1:-1, 1:-1, 1:-1], helm)
        tol = 1e-12 if dt == np.float64 else 1e-5
        if rank == 0:
            print(f"{name}: rel L2 = {err:.3e}", flush=True)
        assert err < tol, f"{name}: {err}"
        worst = max(worst, err / tol)
        halo = np.ones(got.shape, bool)
        halo[1:-1, 1:-1, 1:-1] = False
        assert np.array_equal(got[halo], pl[halo]), "halo cells were modified"
        # z-only solve (solver_gaussel_z) on the same decomposition: slab -> z pencils -> slab, no transforms
        alphai = dt(1.0) / dt(cases.ALPHA)
        bb = (cs["b"] + alphai).astype(dt)
        refz = p.copy()
        O.solver_gaussel_z(ng, ng, ng, cs["a"], bb, cs["c"], cbc[2], cf, alphai, refz)
        pd.copy_(torch.from_numpy(pl))
        cb.solver_gaussel_z(ctx.n, ng, ng, sd.a, torch.from_numpy(bb).to(dev), sd.c, cbc[2], cf, alphai, pd, arrplan=sd.arrplan)
        torch.cuda.synchronize()
        assert ctx.dist_status() == 0, "a device-side barrier timed out"
        gotz = pd.cpu().numpy()
        errz = cases.rel_l2(gotz[1:-1, 1:-1, 1:-1], refz[1 + z0:1 + z1, 1:-1, 1:-1])
        if rank == 0:
            print(f"{name}: solver_gaussel_z rel L2 = {errz:.3e}", flush=True)
        assert errz < tol, f"{name}: gaussel_z {errz}"
        assert np.array_equal(gotz[halo], pl[halo]), "halo cells were modified (gaussel_z)"
        dist.barrier()
        sd.arrplan.destroy()
        ctx.close()
        # distributed TDMA (is_poisson_dtdma, CANSB200_CTX_DTDMA): z stays decomposed, only the reduced system travels.
        # Regular operators only (the reference's gaussel_dtdma has no singular-pivot pin) and nz >= 6 P^2.
        singular = all(bb[0] in "PN" and bb[1] in "PN" for bb in cbc) and not helm
        if not singular and ng[2] >= 6 * world * world and os.environ.get("CANSB200_TEST_DTDMA", "1") == "1":
            ctx2 = cb.Context(ng, is_fp32=dt == np.float32, rank=rank, nranks=world)
            ctx2.connect()
            ctx2.set_dtdma(True)
            assert ctx2.n_z == [ng[0], ng[1], z1 - z0] and ctx2.lo_z == [1, 1, z0 + 1], (ctx2.n_z, ctx2.lo_z)
            sd2 = cb.initsolver(ctx2, ng, cs["dli"], cs["dzci"], cs["dzfi"], cbc, cs["bc"], cf, device=dev)
            assert tuple(sd2.lambdaxy.shape) == (ng[1], ng[0]) and sd2.a.numel() == z1 - z0
            # oracle: the same stages with the distributed elimination on the global field
            ty = dt
            alphai = ty(1.0) / ty(cases.ALPHA)
            px = np.ascontiguousarray(p[1:-1, 1:-1, 1:-1])
            O.fft(cs["arrplan"][0][0], px)
            O.fft(cs["arrplan"][1][0], px)
            q3 = 1 if (cf[2] == "f" and cbc[2][1] == "D") else 0
            bq = (cs["b"] + alphai).astype(dt) if helm else cs["b"]
            nq = ty(cs["normfft"]) * alphai if helm else cs["normfft"]
            O.gaussel_dtdma(dec.zs, ng[2] - q3, cs["a"], bq, cs["c"], cbc[2] == P, nq, px, cs["lambdaxy"])
            O.fft(cs["arrplan"][1][1], px)
            O.fft(cs["arrplan"][0][1], px)
            for rep in range(2):
                pd.copy_(torch.from_numpy(pl))
                if helm:
                    cb.solve_helmholtz(ctx2.n, ng, ctx2.hi(), sd2.arrplan, sd2.normfft, cases.ALPHA, sd2.lambdaxy, sd2.a, sd2.b, sd2.c, None,
                                       None, None, ctx2.is_bound(), cbc, cf, pd)
                else:
                    cb.solver(ctx2.n, ng, sd2.arrplan, sd2.normfft, sd2.lambdaxy, sd2.a, sd2.b, sd2.c, cbc, cf, pd)
            torch.cuda.synchronize()
            assert ctx2.dist_status() == 0, "a device-side barrier timed out (dtdma)"
            gotd = pd.cpu().numpy()
            errd = cases.rel_l2(gotd[1:-1, 1:-1, 1:-1], px[z0:z1])
            if rank == 0:
                print(f"{name}: distributed TDMA rel L2 = {errd:.3e}", flush=True)
            assert errd < tol, f"{name}: dtdma {errd}"
            ndtdma += 1
            assert np.array_equal(gotd[halo], pl[halo]), "halo cells were modified (dtdma)"
            dist.barrier()
            sd2.arrplan.destroy()
            ctx2.close()
    if rank == 0:
        print(f"DIST_OK {worst:.3e} dtdma_cases={ndtdma}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
