import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
HAVE_REFERENCE = os.path.isdir("/root/reference/test")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_linf(a, b):
    """max|a-b| / max|b| -- the tolerance metric of BASELINE.json (1e-3 for the fp32 path)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


@pytest.fixture(scope="session")
def state_dict():
    from rerevst_code_b200.weights import synthetic_state_dict
    from oracle import cases
    return synthetic_state_dict(cases.WEIGHT_SEED)


@pytest.fixture(autouse=True)
def _reset_kernel_tuning():
    """The tensor-core kernel's tuning knobs (rrv_tc_tune*) are process-global: a test that flips them and then fails must not
    leave the library mis-tuned for the tests that follow."""
    yield
    try:
        import torch
        if not torch.cuda.is_available():
            return
        from rerevst_code_b200 import _lib
        if _lib._lib is None:
            return
        lib = _lib._lib
        lib.rrv_tc_tune(256, 2)
        lib.rrv_tc_tune_pair(1, 64)
        lib.rrv_tc_tune_merge(1)
        lib.rrv_tc_tune_pdl(1)
        lib.rrv_set_lo_format(0)
    except Exception:
        pass
