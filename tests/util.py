"""Shared helpers for the test-suite."""
import os

import torch

from conftest import GOLDEN


def golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=True)


def sd_of(fx, prefix="sd."):
    return {k[len(prefix):]: v for k, v in fx.items() if k.startswith(prefix)}


def close(a, b, rtol=1e-4, atol=1e-5, msg=""):
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")
