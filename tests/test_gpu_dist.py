"""Multi-GPU (z-slab) solve: needs >= 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_distributed_solve_matches_oracle(nproc):
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, this box has {torch.cuda.device_count()}: the same kernels, row tables and flags run with "
                    f"virtual ranks on one GPU in tests/test_gpu_dist_local.py, and bench.py --gpus N checks the distributed result "
                    f"against a single-GPU solve on every multi-GPU run")
    port = 29500 + nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    if nproc == 2:   # the distributed-TDMA path (regular operators, nz >= 6 P^2) must have been exercised too
        import re
        m = re.search(r"dtdma_cases=(\d+)", r.stdout)
        assert m and int(m.group(1)) >= 2, r.stdout[-3000:]
