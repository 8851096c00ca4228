"""world_size-2 (and 3) gloo tests of the z-slab decomposition logic on CPU.

The CUDA path stores rows straight into peer memory; what it must reproduce is exactly this data
motion: y-transformed slab -> blocks by destination -> z pencil -> tridiagonal solve with the LOCAL
slice of lambdaxy -> blocks back -> slab.  Here the same index arithmetic (cans_b200/decomp.py, mirrored
in capi.cu) drives an exchange over torch.distributed/gloo with the oracle doing the arithmetic, and the
result must equal the single-rank oracle solve (the reference's transposes are copies on one rank,
dependencies/2decomp-fft/src/transpose_x_to_y.f90:27-36)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("cb_decomp", os.path.join(ROOT, "cans_b200", "decomp.py"))
        decomp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(decomp)   # host logic only: no CUDA library needed in the worker
        import cases
        from oracle import cans_oracle as O
        cs = cases.build_case(name)
        ng, cbc, cf = cs["ng"], cs["cbc"], cs["c_or_f"]
        p = cases.make_rhs(cs)
        ref = cases.oracle_solve(name, cs, p)
        dec = decomp.SlabDecomp(ng, world, rank)
        z0, z1 = dec.z_range()
        y0, y1 = dec.y_range()
        slab = np.ascontiguousarray(p[1 + z0:1 + z1, 1:-1, 1:-1])
        O.fft(cs["arrplan"][0][0], slab)
        O.fft(cs["arrplan"][1][0], slab)
        # forward exchange
        send = [np.ascontiguousarray(b) for b in dec.forward_blocks(slab)]
        allsend = [None] * world
        dist.all_gather_object(allsend, send)
        zp = dec.assemble_zpencil([allsend[s][rank] for s in range(world)])
        assert zp.shape == (ng[2], y1 - y0, ng[0])
        q3 = 1 if (cf[2] == "f" and cbc[2][1] == "D") else 0
        O.gaussel(ng[2] - q3, cs["a"], cs["b"], cs["c"], cbc[2] == ["P", "P"], cs["normfft"], zp, cs["lambdaxy"][y0:y1])
        # way back
        send = [np.ascontiguousarray(b) for b in dec.backward_blocks(zp)]
        dist.all_gather_object(allsend, send)
        slab = dec.assemble_slab([allsend[s][rank] for s in range(world)])
        O.fft(cs["arrplan"][1][1], slab)
        O.fft(cs["arrplan"][0][1], slab)
        err = cases.rel_l2(slab, ref[1 + z0:1 + z1, 1:-1, 1:-1])
        # z-only solve (solver_gaussel_z) over the same exchange, no transforms (src/solver.f90:569-612)
        alphai = 1.0 / cases.ALPHA
        bb = cs["b"] + alphai
        refz = p.copy()
        O.solver_gaussel_z(ng, ng, ng, cs["a"], bb, cs["c"], cbc[2], cf, alphai, refz)
        slab = np.ascontiguousarray(p[1 + z0:1 + z1, 1:-1, 1:-1])
        send = [np.ascontiguousarray(b) for b in dec.forward_blocks(slab)]
        dist.all_gather_object(allsend, send)
        zp = dec.assemble_zpencil([allsend[s][rank] for s in range(world)])
        O.gaussel(ng[2] - q3, cs["a"], bb, cs["c"], cbc[2] == ["P", "P"], alphai, zp, None)
        send = [np.ascontiguousarray(b) for b in dec.backward_blocks(zp)]
        dist.all_gather_object(allsend, send)
        slab = dec.assemble_slab([allsend[s][rank] for s in range(world)])
        err = max(err, cases.rel_l2(slab, refz[1 + z0:1 + z1, 1:-1, 1:-1]))
        # distributed TDMA (is_poisson_dtdma): every rank keeps its slab; only the 2-rows-per-rank reduced system
        # would travel.  The single-process oracle of it must agree with the transposed solve above.
        full = np.ascontiguousarray(p[1:-1, 1:-1, 1:-1])
        lam = cs["lambdaxy"] - 0.37   # regular columns (the pinned mode is the non-distributed solver's business)
        want = full.copy()
        O.gaussel(ng[2] - q3, cs["a"], cs["b"], cs["c"], cbc[2] == ["P", "P"], cs["normfft"], want, lam)
        if all(e - s >= 3 for s, e in zip(dec.zs[:-1], [min(v, ng[2] - q3) for v in dec.zs[1:]])):
            O.gaussel_dtdma(dec.zs, ng[2] - q3, cs["a"], cs["b"], cs["c"], cbc[2] == ["P", "P"], cs["normfft"], full, lam)
            err_d = cases.rel_l2(full[z0:z1], want[z0:z1])
            assert err_d < 1e-9, err_d
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["C3s_channel", "C2s_triperiodic", "odd_sizes", "dirichlet_xyz"])
def test_slab_exchange_reproduces_single_rank_solve(world, name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    errs = dict(q.get(timeout=5) for _ in range(world))
    # same arithmetic, same order along every line: identical up to the order-independent pieces
    assert max(errs.values()) < 1e-14, errs


def test_split_rule_matches_reference_libraries():
    spec_path = os.path.join(ROOT, "cans_b200", "decomp.py")
    import importlib.util
    spec = importlib.util.spec_from_file_location("cb_decomp2", spec_path)
    decomp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(decomp)
    # first n mod P ranks get one extra (decomp_2d.f90:1018-1029)
    assert decomp.split_starts(10, 4) == [0, 3, 6, 8, 10]
    assert decomp.split_starts(512, 8) == list(range(0, 513, 64))
    d = decomp.SlabDecomp([16, 10, 7], 3, 1)
    assert d.n == [16, 10, 2] and d.lo == [1, 1, 4] and d.n_z == [16, 3, 7] and d.lo_z == [1, 5, 1]
