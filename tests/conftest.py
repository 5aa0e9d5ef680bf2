import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle


@pytest.fixture(scope="session")
def ref_backend():
    """The reference's own CUDA backend recompiled for sm_100a (oracle/_ref/, built by oracle/build_ref.py where
    /root/reference is mounted; git-ignored, shipped to the GPU box with the tree).  A GPU run without it FAILS — the parity
    tests would otherwise pass on their golden-file and oracle halves alone and look the same; set RI_ALLOW_NO_REF=1 to turn
    the failure into a (reported) skip."""
    from oracle.build_ref import load_ref
    mod, why = None, "oracle/_ref/_multi_shape_pvcnn_backend.so is missing"
    try:
        mod = load_ref()
    except Exception as e:                       # an unloadable library is as bad as a missing one
        why = "oracle/_ref did not load: %r" % (e,)
    if mod is None:
        if os.environ.get("RI_ALLOW_NO_REF") == "1":
            pytest.skip(why)
        pytest.fail(why + " (build it with `python oracle/build_ref.py` where /root/reference exists)")
    return mod
