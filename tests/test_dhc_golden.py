"""CPU: the oracle against the reference's second known answer, the differentially heated cavity.

`tests/differentially_heated_cavity/test.py:19-20` of the reference holds Nusselt = 8.8252 (rtol = atol = 1e-2) after
10 000 steps of its input.nml.  oracle/dhc_replay.py replays that run around the oracle's solver (REDFT10 / REDFT01 along x,
n = 128: a direction / kind pair the lid-driven cavity does not reach).  The full replay takes ten minutes of one core, so
its result is committed (tests/golden/dhc_nusselt.json, tests/golden/make_dhc_golden.py); this test re-runs the first 40 steps
against that record, checks the recorded Nusselt number against the reference's, and re-runs everything when
CANSB200_LONG_TESTS=1."""
import json
import os

import numpy as np
import pytest

from oracle import cans_oracle as O
from oracle import dhc_replay as D

REC = os.path.join(os.path.dirname(__file__), "golden", "dhc_nusselt.json")


def test_recorded_replay_meets_the_reference_nusselt_number():
    rec = json.load(open(REC))
    assert rec["nstep"] == 10000 and rec["nusselt_ref"] == D.NUSSELT_REF == 8.8252
    np.testing.assert_allclose([rec["nusselt"]], [D.NUSSELT_REF], rtol=1.0e-2, atol=1.0e-2)   # the reference's own bar
    assert rec["divmax"] < O.small(np.float64)


def test_first_steps_reproduce_the_record():
    rec = json.load(open(REC))
    nu, st = D.run_dhc(nstep=40, return_state=True)
    assert nu == pytest.approx(rec["nusselt_after_40_steps"], rel=1e-12)
    assert st["divmax"] < O.small(np.float64)
    # physics of the first steps: the scalar stays between the wall values, buoyancy has started the circulation
    assert st["s"][1:-1, 1:-1, 1:-1].min() >= -0.5 and st["s"][1:-1, 1:-1, 1:-1].max() <= 0.5
    assert np.abs(st["w"]).max() > 1e-3


@pytest.mark.skipif(os.environ.get("CANSB200_LONG_TESTS") != "1", reason="ten minutes of one core: set CANSB200_LONG_TESTS=1")
def test_full_replay_meets_the_reference_nusselt_number():
    nu = D.run_dhc()
    np.testing.assert_allclose([nu], [D.NUSSELT_REF], rtol=1.0e-2, atol=1.0e-2)
