"""Import the UNMODIFIED reference model code (alonet / aloscene) for the model-level benchmark.

Source tree: ``oracle/_ref/aloception_src`` (bundle made by oracle/build_ref_model.py; travels to the GPU box) or, in the
build container, /root/reference.  The reference imports a dozen packages this image lacks (pytorch_lightning, matplotlib,
pycocotools, more_itertools, captum, tensorrt, onnx, ...) at package-import time although the model path uses none of them:
they are satisfied with inert mock modules; ``pkg_resources.get_distribution("aloception")``
(alonet/__init__.py, aloscene/__init__.py) gets a version stub.  The operator is re-pointed with
``aloception_oss_b200.integration.install()``: every ``MSDeformAttn`` of the reference then runs on libmsda_b200.so.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_ROOTS = {"pytorch_lightning", "matplotlib", "pycocotools", "more_itertools", "captum", "tensorrt", "pycuda", "onnx",
              "onnx_graphsurgeon", "onnxsim", "shapely", "nuscenes", "open3d", "pytorch3d", "gdown", "wandb"}


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__, m.__path__, m.__spec__, m.__loader__ = spec.name, [], spec, self
        return m

    def exec_module(self, module):
        pass


class _MockFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in MOCK_ROOTS:
            try:  # only mock what is really absent
                if root not in _MockFinder.checked:
                    _MockFinder.checked[root] = importlib.machinery.PathFinder.find_spec(root) is None
            except Exception:
                _MockFinder.checked[root] = True
            if _MockFinder.checked[root]:
                return importlib.machinery.ModuleSpec(fullname, _MockLoader(), is_package=True)
        return None

    checked = {}


def source_root():
    bundle = os.path.join(ROOT, "oracle", "_ref", "aloception_src")
    if os.path.isfile(os.path.join(bundle, "alonet", "deformable_detr", "deformable_detr_r50.py")):
        return bundle
    ref = os.environ.get("MSDA_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(ref, "alonet")):
        return ref
    return None


_loaded = None


def load():
    """Returns (alonet, aloscene) -- the reference packages, operator re-pointed at the B200 kernels."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = source_root()
    if src is None:
        raise FileNotFoundError("reference model sources not found: run `python oracle/build_ref_model.py` where /root/reference exists")
    if src not in sys.path:
        sys.path.insert(0, src)
    if not any(isinstance(f, _MockFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _MockFinder())
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pkg_resources

        orig = pkg_resources.get_distribution

        def get_distribution(name):
            if name == "aloception":
                return types.SimpleNamespace(version="0.0.0+bundle")
            return orig(name)

        pkg_resources.get_distribution = get_distribution
        import aloscene  # noqa: F401
        import alonet  # noqa: F401
        import alonet.deformable_detr.backbone as ddb
        import alonet.detr.backbone as db

    # never download ImageNet weights (random init: BASELINE.json "random weights")
    db.is_main_process = lambda: False
    ddb.is_main_process = lambda: False
    import aloception_oss_b200.integration as b200

    b200.install()
    _loaded = (alonet, aloscene)
    return _loaded
