#!/usr/bin/env python
"""Generates tests/golden/dhc_nusselt.json: the oracle's replay of the reference's differentially-heated-cavity case
(oracle/dhc_replay.py) over the 10 000 steps the reference's test evaluates (tests/differentially_heated_cavity/test.py:
Nusselt number 8.8252, rtol = atol = 1e-2).  Ten minutes of one CPU core; the CPU suite re-runs the first 40 steps and
checks them against this record, and the whole run when CANSB200_LONG_TESTS=1."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dhc_replay as D  # noqa: E402

t0 = time.time()
nu40 = D.run_dhc(nstep=40)
nu, st = D.run_dhc(nstep=10000, return_state=True)
out = {"case": "tests/differentially_heated_cavity/input.nml (128 x 2 x 128, Ra = 1e6, Pr = 0.71)",
       "nstep": 10000, "nusselt": nu, "nusselt_ref": D.NUSSELT_REF, "rel_err": abs(nu - D.NUSSELT_REF) / D.NUSSELT_REF,
       "nusselt_after_40_steps": nu40, "divmax": st["divmax"], "dt_final": st["dt"], "wall_s": round(time.time() - t0, 1),
       "generated_by": "tests/golden/make_dhc_golden.py"}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "dhc_nusselt.json"), "w"), indent=1)
print(json.dumps(out))
