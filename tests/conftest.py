"""pytest configuration: registers the ``gpu`` marker and puts the repo + product package on sys.path.

``-m "not gpu"``: oracle vs golden vectors / live reference, host logic, host-compiled kernel math,
C-ABI symbol export.  ``-m gpu``: the parity tests proper, through the C ABI on a B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "extended-gan_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def host_harness():
    """ctypes handle of the host build of attn_math.cuh (tests/host_harness)."""
    import ctypes

    d = os.path.join(ROOT, "tests", "host_harness")
    so = os.path.join(d, "libhostharness.so")
    src = os.path.join(d, "harness.cpp")
    hdr = os.path.join(PKG, "csrc", "attn_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)
