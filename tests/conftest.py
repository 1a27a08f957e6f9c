"""Shared test plumbing.

Markers: ``gpu`` = needs a CUDA device (the parity tests proper, run through the
C-ABI on a B200); everything else runs on CPU in the build container.
"""
import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tiny_dense", "tiny_culled", "tiny_rotated", "oneshot_v1", "c1_64x64x32",
                "culled_144x144x24"]
# Token counts of BASELINE configs[2] / configs[4] (1500 / 6000 tokens) through the genuine reference.  Generated
# after the round's last GPU minute: the CPU oracle is pinned to them now; the GPU golden test takes them in once
# they have run on a B200 (the CUDA path is checked against the oracle at these token counts in
# tests/test_gpu_fullsize.py).
GOLDEN_CASES_MANY_TOKENS = ["tokens1500_dense", "tokens6000_culled"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def load_golden(name):
    """-> (frame kwargs dict, S, mode, dict of arrays)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kw = ast.literal_eval(str(z["frame_kwargs"]))
    return kw, int(z["S"]), str(z["mode"]), {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
