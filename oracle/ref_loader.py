"""Import the UNMODIFIED in-tree reference modules (TEST INFRASTRUCTURE, container only).

``/root/reference`` does not exist on the GPU box, so this module is used only by
``oracle/make_golden.py`` and by the CPU tests that skip when the reference is absent.

The reference cannot be imported as-is:
  * ``import ipdb`` (convolutional_gat/baseline_model.py:7, dcgan/model.py:4) -- ipdb is not
    installed; a stub module is placed in ``sys.modules`` for the import.
  * ``self.A = self.A.cuda(h.get_device())`` (baseline_model.py:43,135) raises on CPU tensors;
    ``torch.Tensor.cuda`` is replaced by a no-op *while a reference forward runs*
    (``cpu_shim()`` context manager).
The reference files themselves are never edited or copied.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("CGAT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "convolutional_gat", "baseline_model.py"))


def _load(name: str, relpath: str):
    if "ipdb" not in sys.modules:
        sys.modules["ipdb"] = types.ModuleType("ipdb")
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def baseline_model():
    """The reference's ``convolutional_gat/baseline_model.py`` as a module object."""
    if "bm" not in _cache:
        _cache["bm"] = _load("_ref_baseline_model", "convolutional_gat/baseline_model.py")
    return _cache["bm"]


def dcgan_model():
    """The reference's ``dcgan/model.py`` as a module object."""
    if "dc" not in _cache:
        _cache["dc"] = _load("_ref_dcgan_model", "dcgan/model.py")
    return _cache["dc"]


def kmni_loader():
    """The reference's ``convolutional_gat/data_loaders/kmni_data_loader.py`` as a module object.

    It imports matplotlib / ipdb (absent here; stubbed) and ``..preprocessing.utils`` (loaded from the reference tree
    under the same package names), so the UNMODIFIED ``DataLoader`` class runs on a folder of synthetic files."""
    if "kmni" in _cache:
        return _cache["kmni"]
    for name in ("ipdb", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for pkg in ("_refpkg", "_refpkg.preprocessing", "_refpkg.data_loaders"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    for name, rel in (("_refpkg.preprocessing.utils", "convolutional_gat/preprocessing/utils.py"),
                      ("_refpkg.data_loaders.kmni_data_loader", "convolutional_gat/data_loaders/kmni_data_loader.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    _cache["kmni"] = sys.modules["_refpkg.data_loaders.kmni_data_loader"]
    return _cache["kmni"]


def arai_loader():
    """The reference's ``convolutional_gat/data_loaders/arai_data_loader.py`` as a module object (imports ipdb / tqdm:
    ipdb stubbed); the UNMODIFIED ``DataLoader`` / ``get_loaders`` run on a folder of synthetic files."""
    if "arai" in _cache:
        return _cache["arai"]
    if "ipdb" not in sys.modules:
        sys.modules["ipdb"] = types.ModuleType("ipdb")
    _cache["arai"] = _load("_ref_arai_data_loader", "convolutional_gat/data_loaders/arai_data_loader.py")
    return _cache["arai"]


@contextlib.contextmanager
def cpu_shim():
    """Make ``Tensor.cuda`` a no-op so the reference GAT layers run on CPU tensors."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
