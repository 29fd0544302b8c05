"""The C-ABI library loads and exports every symbol include/cgat_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import PKG, ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "cgat_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgat_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("cgat_attn_fwd", "cgat_attn_bwd", "cgat_conv2d_fprop", "cgat_conv2d_dgrad", "cgat_conv2d_wgrad",
                 "cgat_adj_norm_fwd", "cgat_loss_fwd_bwd", "cgat_adam_step", "cgat_version", "cgat_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    so = os.path.join(PKG, "libcgat_b200.so")
    assert os.path.exists(so), "libcgat_b200.so not built: run __graft_entry__.build()"
    L = ctypes.CDLL(so)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in include/cgat_b200.h but not exported"
    L.cgat_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.cgat_version()


def test_python_binding_covers_header():
    from cgat import _lib

    declared = set(_declared()) - {"cgat_version", "cgat_last_error"}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    _lib.lib()  # loads and sets prototypes


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch

    from cgat.layers import GATMultiHead3D

    layer = GATMultiHead3D(4, 4, 0.2, 1, type_="spatial", n_vertices=6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(torch.rand(1, 4, 4, 4, 6))


def test_product_never_imports_oracle():
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(base, f)
