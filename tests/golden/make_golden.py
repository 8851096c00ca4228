"""Generates tests/golden/*.npz from the oracle (run from the repo root):

    python tests/golden/make_golden.py

Each file holds the solved pressure of one case of tests/cases.py for the
seeded right-hand side `make_rhs(cs, seed=123)`; inputs are regenerated from
the seed, so only the outputs are stored.  The oracle itself is pinned by
tests/test_oracle.py (reference self-test properties) -- these fixtures freeze
its outputs so that the GPU box, which has no /root/reference, checks against
committed numbers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

GOLDEN = ["C1_ldc_2x64x64", "C2s_triperiodic", "C3s_channel", "C4s_duct", "odd_sizes", "dirichlet_xyz",
          "helm_w_face_z", "fp32_channel"]

if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for name in GOLDEN:
        cs = cases.build_case(name)
        p = cases.make_rhs(cs)
        ref = cases.oracle_solve(name, cs, p, helmholtz=name in cases.HELMHOLTZ)
        np.savez_compressed(os.path.join(out, name + ".npz"), p=ref[1:-1, 1:-1, 1:-1])
        print(name, ref.shape, float(np.abs(ref).max()))
