"""Kernel-level parity on the GPU: each hand-written kernel, called through the C ABI (ada_op_*), against the matching
torch op on the same bf16-rounded operands (fp32 torch reference, TF32 off). Cases and tolerances live in
tools/gpu_check.py so the bring-up harness and the test-suite cannot drift apart."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_check  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(gpu_check.CHECKS))
def test_kernel_against_torch(name):
    r = gpu_check.CHECKS[name]()
    assert r["ok"], r
