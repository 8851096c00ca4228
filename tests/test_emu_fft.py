"""Runs the kernels' phase code thread-by-thread on the CPU (tests/emu/emu_fft.cpp): the same
__host__ __device__ functions the CUDA blocks execute, checked against the FFTW definitions."""
import os
import subprocess

import pytest

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "emu_fft")
KINDS = [0, 1, 5, 4, 9, 8]  # R2HC HC2R REDFT10 REDFT01 RODFT10 RODFT01


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [2, 4, 6, 16, 30, 64, 70, 96, 512, 768, 1024])
def test_x_mode(kind, n):
    nl = "7" if n <= 96 else "1"
    r = subprocess.run([EMU, str(n), str(kind), "0", "5", "128", nl, "0", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n,cx", [(2, 16), (22, 8), (64, 16), (512, 16), (768, 16)])
def test_y_mode(kind, n, cx):
    nl = "37" if n <= 96 else "3"
    r = subprocess.run([EMU, str(n), str(kind), "1", str(cx), "256", nl, "0", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("kind", KINDS)
def test_fp32(kind):
    for ymode, tl in ((0, 4), (1, 8)):
        r = subprocess.run([EMU, "256", str(kind), str(ymode), str(tl), "64", "9", "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
