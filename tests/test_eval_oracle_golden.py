"""Pins oracle/eval_oracle.py (validation-loop post-ops) against the values the reference's own
src/util/alignment.py and src/util/metric.py produced (tests/golden/eval/, make_golden_eval.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import eval_oracle as EO

EVAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval")


def golden_cases():
    return sorted(glob.glob(os.path.join(EVAL, "*.npz")))


def load_case(path):
    z = np.load(path)
    s = EO.synth_sample(int(z["seed"]), int(z["h"]), int(z["w"]), int(z["H"]), int(z["W"]))
    want = {"scale": float(z["scale"]), "shift": float(z["shift"]),
            "pred": {m: float(z["pred_" + m]) for m in EO.METRICS}, "aligned": {m: float(z["aligned_" + m]) for m in EO.METRICS}}
    return s, want


@pytest.mark.parametrize("path", golden_cases())
def test_eval_oracle_matches_reference_functions(path):
    s, want = load_case(path)
    got = EO.evaluate_sample(**s)
    assert abs(got["scale"] - want["scale"]) < 1e-6 and abs(got["shift"] - want["shift"]) < 1e-6
    for v in ("pred", "aligned"):
        for m in EO.METRICS:
            assert abs(got[v][m] - want[v][m]) <= 1e-6 * max(1.0, abs(want[v][m])), (v, m)
