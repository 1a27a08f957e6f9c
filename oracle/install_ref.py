"""TEST / BENCH INFRASTRUCTURE ONLY -- make the genuine reference travel to the GPU box.

``/root/reference`` exists only in the build container; ``gpurun`` ships ``/root/repo``.  This script copies the
parts of the reference tree its hot path needs (``lib/``, ``kmeans_dict/``, ``configs/`` and the chumpy-free
SMPL pickle, ~40 MB) UNMODIFIED into ``baseline/_ref/`` -- git-ignored (never part of the history, never product
source) but not gpurun-ignored -- so that on the B200

  * ``bench.py`` can time the reference's own ``Renderer.render`` in torch-CUDA (TF32 off and on) beside the
    fused path, and its CPU path through the reference's own code instead of the oracle port,
  * ``tests/test_gpu_reference_cuda.py`` can compare full frames with the reference itself, and resolve the
    plugin through the reference's ``make_renderer``.

The reference is not a Python package (no setup.py / pyproject), so the ``pip install --target baseline/_ref``
recipe does not apply.  ``oracle/ref_shim.py`` finds the copy when ``/root/reference`` is absent.
Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("TRANSHUMAN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ["lib", "kmeans_dict", "configs",
         os.path.join("third_parties", "smpl", "models", "basicModel_neutral_lbs_10_207_0_v1.0.0.pkl")]


def install(force: bool = False) -> bool:
    """-> True if baseline/_ref is in place afterwards."""
    if not os.path.isdir(os.path.join(SRC, "lib", "networks")):
        return os.path.isdir(os.path.join(DST, "lib", "networks"))
    stamp = os.path.join(DST, ".installed")
    if os.path.exists(stamp) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    with open(stamp, "w") as f:
        f.write("copied unmodified from %s by oracle/install_ref.py\n" % SRC)
    return True


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "ready" if ok else "reference tree not available")
