"""Host logic of the conv dispatch (no GPU, no compute calls): which shapes `cgat_conv_tc_supported` hands to the tcgen05
kernels -- streamed operands from 64 channels up (csrc/conv_tc_big.cu), resident weights below (csrc/conv_tc.cu) -- and
which stay on the CUDA-core kernels.  The shapes are the reference's own: the conv-GAT node conv, the DCGAN nets
(dcgan/model.py:55-179, discriminator convs after the 2x2 regrouping of cgat.conv_layers) and SmaAt-UNet convs."""
import ctypes

import pytest

from cgat import _lib
from cgat.functional import _conv_desc


def _support(n, h, w, cin, cout, k, pad, ho, wo, dtype=None, stride=1, groups=1):
    d = _conv_desc(n, h, w, cin, cout, k, k, stride, pad, pad, ho, wo, _lib.BF16 if dtype is None else dtype, 0, groups)
    L = _lib.lib()
    return [L.cgat_conv_tc_supported(ctypes.byref(d), i) for i in range(3)], d


@pytest.mark.parametrize("shape, expect", [
    ((64, 64, 64, 24, 72, 3, 1, 64, 64), [1, 1, 1]),      # conv-GAT block-diagonal node conv: resident-weight kernels
    ((64, 17, 17, 256, 128, 2, 0, 16, 16), [1, 1, 1]),    # FD/TD conv2 regrouped: streamed kernels
    ((64, 9, 9, 512, 256, 2, 0, 8, 8), [1, 1, 1]),        # conv3
    ((64, 5, 5, 1024, 512, 2, 0, 4, 4), [1, 1, 1]),       # conv4
    ((64, 33, 33, 16, 64, 2, 0, 32, 32), [1, 1, 1]),      # conv1 regrouped (16 channels): resident-weight kernels
    ((2, 16, 16, 2048, 512, 1, 0, 16, 16), [1, 1, 1]),    # SmaAt-UNet pointwise
    ((64, 64, 64, 32, 16, 4, 1, 64, 64), [1, 1, 0]),      # generator layer: wgrad on the small-channel CUDA-core kernel
    ((2, 16, 16, 32, 16, 4, 1, 16, 16), [1, 1, 1]),       # same layer on few pixels: wgrad on the streamed kernel
    ((64, 64, 64, 4, 32, 4, 1, 64, 64), [0, 1, 0]),       # cin = 4: no 16-byte channel rows for TMA
    ((64, 4, 4, 512, 1, 4, 0, 1, 1), [0, 0, 0]),          # discriminators' last conv (cout = 1): full-window dot kernel
])
def test_tensor_core_support_matrix(shape, expect):
    got, _ = _support(*shape)
    assert got == expect


def test_never_on_tensor_cores():
    assert _support(2, 16, 16, 64, 64, 3, 1, 16, 16, dtype=_lib.F32)[0] == [0, 0, 0]      # fp32 stays exact on CUDA cores
    assert _support(2, 16, 16, 64, 128, 4, 1, 8, 8, stride=2)[0] == [0, 0, 0]             # stride 2: regrouped by the module
    assert _support(2, 16, 16, 64, 128, 3, 1, 16, 16, groups=64)[0] == [0, 0, 0]          # depthwise: conv_depthwise.cu


def test_workspace_sizes():
    L = _lib.lib()
    _, d = _support(64, 9, 9, 512, 256, 2, 0, 8, 8)
    assert L.cgat_conv_workspace_bytes(ctypes.byref(d), 0) == 0                       # fprop reads the KRSC weights directly
    assert L.cgat_conv_workspace_bytes(ctypes.byref(d), 1) == 2 * 2 * 512 * 256 * 2   # dgrad: rotated + transposed bf16 copy
