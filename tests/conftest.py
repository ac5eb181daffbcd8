import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


def maxnorm_err(a, b):
    """The parity metric: max|a-b| / max|b| per tensor."""
    import numpy as np
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = np.abs(b).max()
    if denom == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / denom)


def record(name, value):
    """Append one measured parity figure to gpurun_out/parity_measured.jsonl (scratch; the GPU box brings it back) so
    that the numbers behind the asserted bounds can be quoted in DESIGN.md / profiles/."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, "value": value}) + "\n")
    except OSError:
        pass
    print("[measured] %s = %s" % (name, value))
