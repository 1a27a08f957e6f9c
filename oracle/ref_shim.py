"""TEST INFRASTRUCTURE ONLY -- import the genuine TransHuman reference on CPU.

This module is *not* part of the product path.  It exists so that
``oracle/make_golden.py`` and the ``-m "not gpu"`` tests can execute the
reference's own ``Network`` / ``Renderer`` / ``raw2outputs`` from
``/root/reference`` (read-only, only present in the build container) in order
to pin ``oracle/transhuman_oracle.py`` and to generate the committed fixtures
under ``tests/golden/``.  Nothing here is reachable from ``transhuman_b200``.

What blocks a plain ``import`` of the reference, and what the shim does about
it:

* ``lib/config/config.py:1`` imports open3d and runs argparse at import time
  (``config.py:152-167``)            -> stub module + preset ``sys.argv``;
* ``lib/networks/make_network.py:2`` imports ``imp`` (gone in Python 3.12)
                                      -> stub module;
* ``cross_transformer.py:5-14,29`` import spconv and pytorch3d
                                      -> stub modules, ``knn_points`` injected
                                         from the oracle's fully specified
                                         restatement;
* ``if_clight_renderer.py:6-26`` import matplotlib, chumpy, trimesh, open3d
                                      -> stub modules;
* ``Renderer.__init__`` (``if_clight_renderer.py:43,55``) opens
  ``./data/smplx/smpl/SMPL_NEUTRAL.pkl`` and ``./kmeans_dict/...`` relative to
  the CWD                             -> scratch CWD populated by the caller
                                         (real files or synthetic ones);
* ``.cuda()`` / ``torch.cuda.current_device()`` at
  ``if_clight_renderer.py:180-181,195`` -> identity on CPU.

No reference file is modified or copied; modules are loaded in place with
``importlib`` (the ``imp.load_source`` equivalent).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

def _find_reference_root() -> str:
    """``$TRANSHUMAN_REFERENCE``, else ``/root/reference`` (build container), else the unmodified copy that
    ``oracle/install_ref.py`` ships to the GPU box under the git-ignored ``baseline/_ref/``."""
    env = os.environ.get("TRANSHUMAN_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/lib/networks"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REFERENCE_ROOT = _find_reference_root()
_STATE: dict = {}
_STUBBED: set = set()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "networks"))


class _Anything:
    """Attribute sink: any attribute / call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)

    def _getattr(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything

    mod.__getattr__ = _getattr  # type: ignore[attr-defined]
    mod.__path__ = []  # behave like a package so sub-imports resolve
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    sys.modules[name] = mod
    return mod


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolve any ``stubbed_pkg.sub.module`` import to an attribute sink."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _stub(spec.name)

    def exec_module(self, module):
        pass


def _install_stubs(knn_points):
    import torch.nn as nn

    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())

    class _NoopModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, x):
            return x

    class _SparseSequential(nn.Sequential):
        def __init__(self, *mods, **k):
            super().__init__(*[m for m in mods if isinstance(m, nn.Module)])

    for name in [
        "open3d", "matplotlib", "matplotlib.pyplot", "mpl_toolkits",
        "mpl_toolkits.mplot3d", "chumpy", "trimesh", "mcubes", "imageio",
        "termcolor", "plyfile", "sklearn", "sklearn.neighbors", "sklearn.cluster",
        "skimage", "skimage.measure", "skimage.metrics", "lpips", "tensorboardX", "easydict",
        "imgaug", "imgaug.augmenters", "spconv", "spconv.pytorch",
        "spconv.pytorch.core", "spconv.pytorch.identity", "spconv.pytorch.ops",
        "spconv.pytorch.pool", "spconv.pytorch.tables", "spconv.pytorch.utils",
        "pytorch3d", "pytorch3d.structures", "pytorch3d.renderer",
    ]:
        top = name.split(".")[0]
        if top in _STUBBED or (top not in sys.modules and importlib.util.find_spec(top) is None):
            _STUBBED.add(top)
            _stub(name)
    conv_names = ["SparseConv2d", "SparseConv3d", "SparseConvTranspose2d",
                  "SparseConvTranspose3d", "SparseInverseConv2d",
                  "SparseInverseConv3d", "SubMConv2d", "SubMConv3d"]
    _stub("spconv.pytorch.conv", **{n: _NoopModule for n in conv_names})
    _stub("spconv.pytorch.modules", SparseModule=_NoopModule,
          SparseSequential=_SparseSequential)
    _stub("pytorch3d.ops", knn_points=knn_points)
    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")

        def load_source(module, path):
            spec = importlib.util.spec_from_file_location(module, path)
            m = importlib.util.module_from_spec(spec)
            sys.modules[module] = m
            spec.loader.exec_module(m)
            return m

        imp.load_source = load_source
        sys.modules["imp"] = imp




def make_scratch_cwd(smpl_pkl: dict | None = None, kmeans: dict | None = None) -> str:
    """Create the scratch CWD the reference expects.

    ``smpl_pkl`` : dict with ``v_template`` (6890,3) f64 and ``f`` -- written as
                   ``data/smplx/smpl/SMPL_NEUTRAL.pkl``; ``None`` links the
                   reference's own chumpy-free SMPL pickle.
    ``kmeans``   : ``{num_class: pc2voxel_ind int array}`` -- each written in
                   the reference's ``kmeans_dict_{n}.npy`` format; ``None``
                   links the reference's ``kmeans_dict`` directory.
    """
    d = tempfile.mkdtemp(prefix="th_refcwd_")
    for sub in ("lib", "configs", "third_parties"):
        os.symlink(os.path.join(REFERENCE_ROOT, sub), os.path.join(d, sub))
    os.makedirs(os.path.join(d, "data", "smplx", "smpl"))
    pkl = os.path.join(d, "data", "smplx", "smpl", "SMPL_NEUTRAL.pkl")
    if smpl_pkl is None:
        os.symlink(os.path.join(REFERENCE_ROOT, "third_parties", "smpl", "models",
                                "basicModel_neutral_lbs_10_207_0_v1.0.0.pkl"), pkl)
    else:
        with open(pkl, "wb") as f:
            pickle.dump(smpl_pkl, f)
    if kmeans is None:
        os.symlink(os.path.join(REFERENCE_ROOT, "kmeans_dict"), os.path.join(d, "kmeans_dict"))
    else:
        os.makedirs(os.path.join(d, "kmeans_dict"))
        for n, pc2voxel in kmeans.items():
            pc2voxel = np.asarray(pc2voxel, dtype=np.int32)
            v2pc = {np.int32(c): [np.int32(i) for i in np.nonzero(pc2voxel == c)[0]]
                    for c in range(int(pc2voxel.max()) + 1)}
            np.save(os.path.join(d, "kmeans_dict", f"kmeans_dict_{n}.npy"),
                    {"pc2voxel_ind": pc2voxel, "dict_voxel2pc_ind": v2pc},
                    allow_pickle=True)
    return d


def load_reference(knn_points, cwd: str, opts: dict | None = None, device: str = "cpu"):
    """Import the reference modules.  Returns a namespace with
    ``cfg, cross_transformer, renderer_mod, mesh_renderer_mod, nerf_net_utils,
    embedder, vision_transformer``.  Can only be done once per process (the
    reference keeps a global ``cfg``); later calls update ``cfg`` in place.
    ``device="cuda"`` leaves ``.cuda()`` alone (the reference in torch-CUDA on the GPU box); the default makes
    it an identity so the CUDA-only prologue runs on the CPU."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    os.chdir(cwd)
    if "ns" in _STATE:
        ns = _STATE["ns"]
        sys.modules["pytorch3d.ops"].knn_points = knn_points
        ns.cross_transformer.knn_points = knn_points
        ns.renderer_mod.knn_points = knn_points
        ns.mesh_renderer_mod.knn_points = knn_points
        for k, v in (opts or {}).items():
            setattr(ns.cfg, k, v)
        return ns
    _install_stubs(knn_points)
    if cwd not in sys.path:
        sys.path.insert(0, cwd)
    argv = sys.argv
    sys.argv = ["ref_shim", "--cfg_file", "configs/train_or_eval.yaml",
                "pretrained", "False", "gpus", "[0]"]
    try:
        if device == "cpu":
            torch.Tensor.cuda = lambda self, *a, **k: self  # CPU only
            torch.cuda.current_device = lambda: "cpu"
        cfg = importlib.import_module("lib.config").cfg
    finally:
        sys.argv = argv
    for k, v in (opts or {}).items():
        setattr(cfg, k, v)
    ns = types.SimpleNamespace(cfg=cfg)
    ns.cross_transformer = importlib.import_module("lib.networks.cross_transformer")
    ns.renderer_mod = importlib.import_module("lib.networks.renderer.if_clight_renderer")
    ns.mesh_renderer_mod = importlib.import_module("lib.networks.renderer.if_mesh_renderer")
    ns.nerf_net_utils = importlib.import_module("lib.networks.renderer.nerf_net_utils")
    ns.embedder = importlib.import_module("lib.networks.embedder")
    ns.vision_transformer = importlib.import_module("lib.networks.vision_transformer")
    ns.make_renderer = importlib.import_module("lib.networks.renderer.make_renderer")
    _STATE["ns"] = ns
    return ns
