"""``nn.Conv2d`` drop-in on the CUDA conv kernels (K1/K2/K3) for the conv stacks around the GAT layer.

API and ``state_dict`` layout are those of ``torch.nn.Conv2d`` (``weight [cout, cin/groups, kh, kw]``, ``bias [cout]``),
so the DCGAN nets (dcgan/model.py:19-179) and SmaAt-UNet keep their checkpoint keys.  Tensors keep PyTorch's
NCHW *shape* but live in channels_last memory, which is exactly the NHWC layout the kernels take, so a stack of
these layers (with BatchNorm / pooling from PyTorch in between) never transposes.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from .functional import IMPL_AUTO, conv2d_nhwc


class _S2DPad(torch.autograd.Function):
    """``x [N, H, W, C]`` -> zero-padded by one pixel and regrouped in 2x2 blocks ``[N, H/2+1, W/2+1, 4C]``; the backward is
    the inverse gather (csrc/layout_kernels.cu)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, C = x.shape
        xs = torch.empty(N, H // 2 + 1, W // 2 + 1, 4 * C, device=x.device, dtype=x.dtype)
        _lib.call("cgat_s2d_pad", _lib.ptr(x), _lib.ptr(xs), _lib.dtype_tag(x), N, H, W, C, 0, _lib.stream())
        ctx.shape = (N, H, W, C)
        return xs

    @staticmethod
    def backward(ctx, dxs):
        N, H, W, C = ctx.shape
        dxs = dxs.contiguous()
        dx = torch.empty(N, H, W, C, device=dxs.device, dtype=dxs.dtype)
        _lib.call("cgat_s2d_pad", _lib.ptr(dxs), _lib.ptr(dx), _lib.dtype_tag(dxs), N, H, W, C, 1, _lib.stream())
        return dx


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, groups=1, act=0,
                 impl=IMPL_AUTO):
        super().__init__()
        kh, kw = _pair(kernel_size)
        sh, sw = _pair(stride)
        if sh != sw:
            raise ValueError("only equal strides are supported")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, (kh, kw)
        self.stride, self.groups, self.act, self.impl = sh, groups, act, impl
        if padding == "same":
            if sh != 1:
                raise ValueError("padding='same' needs stride 1")
            # PyTorch: total k-1, the extra element of an even kernel goes AFTER (left 1 / right 2 for k=4)
            th, tw = kh - 1, kw - 1
            self.pad = (th // 2, tw // 2, th - th // 2, tw - tw // 2)  # top, left, bottom, right
        else:
            ph, pw = _pair(padding)
            self.pad = (ph, pw, ph, pw)
        self.padding = padding
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, kh, kw))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # same init as torch.nn.Conv2d
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[1] * self.weight.shape[2] * self.weight.shape[3]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def _space_to_depth_route(self, x_nhwc):
        """The DCGAN discriminators' ``k=4, stride=2, padding=1`` convs (dcgan/model.py:150-165) as a stride-1 2x2 conv
        over the zero-padded input regrouped in 2x2 pixel blocks (4*cin channels): window rows ``2i-1 .. 2i+2`` are the
        blocks ``i, i+1``, so nothing is wasted and the tcgen05 implicit-GEMM kernels (stride 1 only) serve it.  Returns
        ``None`` when the shape is not of that class or the tensor-core path would not take the regrouped conv."""
        import ctypes

        N, H, W, C = x_nhwc.shape
        if not (self.stride == 2 and self.kernel_size == (4, 4) and self.pad == (1, 1, 1, 1) and self.groups == 1
                and self.impl == IMPL_AUTO and x_nhwc.is_cuda and x_nhwc.dtype == torch.bfloat16 and H % 2 == 0
                and W % 2 == 0 and (4 * C) % 8 == 0):
            return None
        hs, ws = H // 2 + 1, W // 2 + 1
        d = _lib.ConvDesc(N, hs, ws, 4 * C, self.out_channels, 2, 2, 1, 0, 0, hs - 1, ws - 1, _lib.BF16, self.act, 1)
        if not _lib.lib().cgat_conv_tc_supported(ctypes.byref(d), 0):
            return None
        xs = _S2DPad.apply(x_nhwc)  # [N, hs, ws, 4C], channel = (a, b, c): pad + regroup in one pass (cgat_s2d_pad)
        # weight[cout, c, 2r+a, 2s+b] -> [cout, r, s, (a, b, c)]
        w = self.weight.view(self.out_channels, C, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(self.out_channels, 2, 2, 4 * C)
        return conv2d_nhwc(xs, w, self.bias, stride=1, pad=(0, 0, 0, 0), act=self.act, impl=IMPL_AUTO)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x[N, C, H, W]`` (any memory format) -> ``[N, C', H', W']`` in channels_last memory."""
        x_nhwc = x.permute(0, 2, 3, 1)  # a view; contiguous iff x is channels_last
        y = self._space_to_depth_route(x_nhwc)
        if y is not None:
            return y.permute(0, 3, 1, 2)
        w_krsc = self.weight.permute(0, 2, 3, 1)
        y = conv2d_nhwc(x_nhwc, w_krsc, self.bias, stride=self.stride, pad=self.pad, act=self.act, impl=self.impl,
                        groups=self.groups)
        return y.permute(0, 3, 1, 2)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, groups={self.groups}")
