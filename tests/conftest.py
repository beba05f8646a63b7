import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    config.addinivalue_line("markers", "slow: minutes on a B200 (full-length BASELINE configurations)")


def pytest_collection_modifyitems(config, items):
    import torch
    from oracle import ref_loader
    has_gpu = torch.cuda.is_available()
    has_ref = ref_loader.reference_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


def observed(name, value, bound):
    """Asserts value < bound and prints the observed figure (pytest -s / -rP shows it): the tolerances written in the tests are
    kept at a small multiple of what is measured, and this is where the measurement comes from."""
    print("OBSERVED %s = %.3g (bound %.3g)" % (name, value, bound))
    assert value < bound, "%s = %.3g exceeds %.3g" % (name, value, bound)
