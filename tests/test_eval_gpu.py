"""ada_eval_sample (validation-loop post-ops on the device) on a B200 against the fixtures
minted from the reference's own alignment / metric functions and against the oracle. Bars: scale/shift 1e-4 relative (the
reference solves the least squares in fp32 with LAPACK, the kernel with fp64 normal equations), metrics 1e-4 relative
(fp32 per-pixel terms, fp64 vs torch's fp32 summation), threshold accuracies 1e-5 absolute."""
import pytest
import torch

from amodal_depth_anything_b200 import ops
from oracle import eval_oracle as EO
from tests.test_eval_oracle_golden import golden_cases, load_case

pytestmark = pytest.mark.gpu


def _check(got, want):
    assert abs(got["scale"] - want["scale"]) <= 1e-4 * abs(want["scale"])
    assert abs(got["shift"] - want["shift"]) <= 1e-4 * max(abs(want["shift"]), 0.1)
    for v in ("pred", "aligned"):
        for m in EO.METRICS:
            tol = 1e-5 if m.startswith("delta") else 1e-4 * max(abs(want[v][m]), 1e-3)
            if v == "aligned":
                tol *= 3  # inherits the 1e-4 of scale/shift
            assert abs(got[v][m] - want[v][m]) <= tol, (v, m, got[v][m], want[v][m])


@pytest.mark.parametrize("path", golden_cases())
def test_eval_sample_matches_reference_functions(path):
    s, want = load_case(path)
    out = ops.eval_sample(*[s[k].cuda() for k in ("pred", "depth_gt", "depth_obs", "visible_mask", "object_mask")])
    got = ops.eval_sample_dict(out)
    assert got["n_visible"] == int(s["visible_mask"].sum()) and got["n_object"] == int(s["object_mask"].sum())
    _check(got, want)
    _check(got, EO.evaluate_sample(**s))


def test_eval_sample_nan_semantics():
    """A non-positive aligned depth makes the log metrics NaN in the reference (the trainer then skips the sample,
    discriminative_trainer.py:596-597); the kernel reports NaN for the same metrics and finite values for the others."""
    s = EO.synth_sample(64, 70, 70, 120, 160)
    s["depth_obs"] = 1.0 - 2.0 * s["depth_gt"]            # alignment maps part of the object region below zero
    want = EO.evaluate_sample(**s)
    got = ops.eval_sample_dict(ops.eval_sample(*[s[k].cuda() for k in ("pred", "depth_gt", "depth_obs", "visible_mask", "object_mask")]))
    for m in EO.METRICS:
        a, b = got["aligned"][m], want["aligned"][m]
        assert (a != a) == (b != b), (m, a, b)
        if b == b:  # 1/o next to a zero crossing amplifies the 1e-4 of scale/shift: i_rmse gets a looser bar here
            assert abs(a - b) <= (5e-2 if m == "i_rmse" else 1e-3) * max(abs(b), 1e-3), (m, a, b)
