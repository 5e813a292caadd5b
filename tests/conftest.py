import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


from oracle.model_export import oracle_model_from_export, random_inputs  # noqa: E402,F401  (re-exported for the tests)


def rel_err(a, ref):
    """max |a - ref| / max(|ref|_inf, 1): relative to the scale of the reference array."""
    a, ref = np.asarray(a), np.asarray(ref)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1.0))


def rel_err_rows(a, ref, axis=-1, floor=1e-6):
    """Worst per-plane relative error: every component plane (all leading indices; the last axis runs over the units) is scaled
    by its OWN magnitude max_u |ref|, with the explicit floor eps * scale = 1e-6 * max |ref| (or 1e-6 when the whole array is
    below 1), so an entry 10^6 times smaller than the largest block entry is still checked to 1e-9 of its own plane
    ("relative to max(|ref|, eps * scale)")."""
    a, ref = np.asarray(a), np.asarray(ref)
    scale = np.maximum(np.abs(ref).max(axis=axis, keepdims=True), floor * max(np.abs(ref).max(), 1.0))
    return float((np.abs(a - ref) / scale).max())
