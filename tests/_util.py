import os

import numpy as np

TOL = 1e-5     # north_star: fp32 results within 1e-5 relative


def load_golden(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        import pytest
        pytest.skip("golden fixture %s not generated yet" % name)
    return np.load(p)


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) elementwise-max — the parity metric of SURVEY.md §8c."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))) if a.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| — for sums whose individual terms cancel (voxel means, gradients via float atomics)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if a.size else 0.0
