"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs in the CPU-only authoring container (oracle vs golden, host
logic, C-ABI symbol checks); ``-m gpu`` runs on a B200 and compares the CUDA path,
called through the C-ABI, with the oracle.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(GOLDEN / name)
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else {}
    return z, meta


def golden_state(z, prefix):
    import torch
    out = {}
    for k in z.files:
        if k.startswith(prefix + "/"):
            out[k[len(prefix) + 1:]] = torch.from_numpy(np.array(z[k]))
    return out


MODEL_CASES = ["c6_t640", "c3_t336_ternary", "c14_t3840", "c8_h32_l1", "c6_t3840_b8"]
