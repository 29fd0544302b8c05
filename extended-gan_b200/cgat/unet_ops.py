"""The non-conv ops of the SmaAt-UNet (convolutional_gat/unet_model.py:20-29 applies the net per vertex) on the CUDA kernels
of ``csrc/unet_glue_kernels.cu``: 2x2 max-pooling, bilinear x2 up-sampling + pad + concat, and CBAM's channel and spatial
gates, each with its backward.  Modules mirror the ``torch.nn`` ones they replace (same sub-module names and ``state_dict``
keys inside SmaAt_UNet.py); tensors keep PyTorch's NCHW shape in channels_last memory = the kernels' NHWC.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, ptr, require_cuda, stream


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous()


def _check(x):
    require_cuda(x)
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"the UNet ops serve fp32 and bf16 activations, got {x.dtype}")


class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _check(x)
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        y = torch.empty(N, H // 2, W // 2, C, device=x.device, dtype=x.dtype)
        idx = torch.empty(N, H // 2, W // 2, C, device=x.device, dtype=torch.uint8)
        _lib.call("cgat_maxpool2_fwd", ptr(xh), ptr(y), ptr(idx), _lib.dtype_tag(xh), N, H, W, C, stream())
        ctx.save_for_backward(idx)
        ctx.shape = (N, H, W, C)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, H, W, C = ctx.shape
        dyh = _nhwc(dy)
        dx = torch.empty(N, H, W, C, device=dy.device, dtype=dy.dtype)
        _lib.call("cgat_maxpool2_bwd", ptr(dyh), ptr(idx), ptr(dx), _lib.dtype_tag(dyh), N, H, W, C, stream())
        return dx.permute(0, 3, 1, 2)


class MaxPool2d(nn.Module):
    """``nn.MaxPool2d(2)``."""

    def forward(self, x):
        return _MaxPool2.apply(x)


class _UpCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2):
        _check(x1)
        _check(x2)
        a, b = _nhwc(x1), _nhwc(x2).to(x1.dtype)
        N, h1, w1, c1 = a.shape
        _, H, W, c2 = b.shape
        out = torch.empty(N, H, W, c2 + c1, device=x1.device, dtype=x1.dtype)
        _lib.call("cgat_upcat_fwd", ptr(a), ptr(b), ptr(out), _lib.dtype_tag(a), N, h1, w1, c1, H, W, c2, stream())
        ctx.shape = (N, h1, w1, c1, H, W, c2)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        N, h1, w1, c1, H, W, c2 = ctx.shape
        d = _nhwc(dout)
        dx1 = torch.empty(N, h1, w1, c1, device=d.device, dtype=d.dtype)
        dx2 = torch.empty(N, H, W, c2, device=d.device, dtype=d.dtype)
        _lib.call("cgat_upcat_bwd", ptr(d), ptr(dx1), ptr(dx2), _lib.dtype_tag(d), N, h1, w1, c1, H, W, c2, stream())
        return dx1.permute(0, 3, 1, 2), dx2.permute(0, 3, 1, 2)


def upsample_pad_concat(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """``cat([x2, pad(Upsample(scale_factor=2, mode="bilinear", align_corners=True)(x1))], dim=1)`` (UpDS.forward) in one
    pass: the up-sampled and the padded tensor are never materialised."""
    return _UpCat.apply(x1, x2)


class _ChannelGate(torch.autograd.Function):
    """CBAM channel attention: ``x * sigmoid(MLP(avgpool(x)) + MLP(maxpool(x)))``."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        _check(x)
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        hid = w1.shape[0]
        dev, dt = x.device, _lib.dtype_tag(xh)
        f32 = dict(device=dev, dtype=torch.float32)
        ws = torch.empty(lib().cgat_pool_hw_workspace_bytes(N, H * W, C), device=dev, dtype=torch.uint8)
        avg, mx = torch.empty(N, C, **f32), torch.empty(N, C, **f32)
        arg = torch.empty(N, C, device=dev, dtype=torch.int32)
        st = stream()
        _lib.call("cgat_pool_hw", ptr(xh), dt, N, H * W, C, ptr(ws), ptr(avg), ptr(mx), ptr(arg), st)
        w1c, b1c, w2c, b2c = (t.detach().float().contiguous() for t in (w1, b1, w2, b2))
        pre, scale = torch.empty(N, 2, hid, **f32), torch.empty(N, C, **f32)
        _lib.call("cgat_cbam_mlp_fwd", ptr(avg), ptr(mx), ptr(w1c), ptr(b1c), ptr(w2c), ptr(b2c), N, C, hid, ptr(pre),
                  ptr(scale), st)
        y = torch.empty_like(xh)
        _lib.call("cgat_gate_channels_fwd", ptr(xh), ptr(scale), ptr(y), dt, N, H * W, C, st)
        ctx.save_for_backward(xh, avg, mx, arg, pre, scale, w1c, w2c, ws)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, avg, mx, arg, pre, scale, w1c, w2c, ws = ctx.saved_tensors
        N, H, W, C = xh.shape
        hid = w1c.shape[0]
        dyh = _nhwc(dy).to(xh.dtype)
        dt, st = _lib.dtype_tag(xh), stream()
        f32 = dict(device=xh.device, dtype=torch.float32)
        dscale = torch.empty(N, C, **f32)
        _lib.call("cgat_dot_hw", ptr(xh), ptr(dyh), dt, N, H * W, C, ptr(ws), ptr(dscale), st)
        ds, dpre = torch.empty(N, C, **f32), torch.empty(N, 2, hid, **f32)
        davg, dmax = torch.empty(N, C, **f32), torch.empty(N, C, **f32)
        dw1, db1 = torch.empty_like(w1c), torch.empty(hid, **f32)
        dw2, db2 = torch.empty_like(w2c), torch.empty(C, **f32)
        _lib.call("cgat_cbam_mlp_bwd", ptr(dscale), ptr(scale), ptr(pre), ptr(avg), ptr(mx), ptr(w1c), ptr(w2c), N, C, hid,
                  ptr(ds), ptr(dpre), ptr(davg), ptr(dmax), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), st, launches=2)
        dx = torch.empty_like(xh)
        _lib.call("cgat_gate_channels_bwd", ptr(dyh), ptr(scale), ptr(davg), ptr(dmax), ptr(arg), ptr(dx), dt, N, H * W, C, st)
        return dx.permute(0, 3, 1, 2), dw1, db1, dw2, db2


class _ChanPool(torch.autograd.Function):
    """``cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)`` (CBAM spatial attention's input)."""

    @staticmethod
    def forward(ctx, x):
        _check(x)
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        o = torch.empty(N, H, W, 2, device=x.device, dtype=x.dtype)
        arg = torch.empty(N, H, W, device=x.device, dtype=torch.int32)
        _lib.call("cgat_chan_pool_fwd", ptr(xh), ptr(o), ptr(arg), _lib.dtype_tag(xh), N * H * W, C, stream())
        ctx.save_for_backward(arg)
        ctx.c = C
        return o.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, do):
        (arg,) = ctx.saved_tensors
        N, H, W = arg.shape
        d = _nhwc(do)
        dx = torch.empty(N, H, W, ctx.c, device=d.device, dtype=d.dtype)
        _lib.call("cgat_chan_pool_bwd", ptr(d), ptr(arg), ptr(dx), _lib.dtype_tag(d), N * H * W, ctx.c, stream())
        return dx.permute(0, 3, 1, 2)


class _PixelGate(torch.autograd.Function):
    """``x * s`` with ``s [N, 1, H, W]``."""

    @staticmethod
    def forward(ctx, x, s):
        _check(x)
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        sh = s.reshape(N, H, W).to(x.dtype).contiguous()
        y = torch.empty_like(xh)
        _lib.call("cgat_gate_pixels", ptr(xh), ptr(sh), ptr(y), _lib.dtype_tag(xh), N, H * W, C, stream())
        ctx.save_for_backward(xh, sh)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, sh = ctx.saved_tensors
        N, H, W, C = xh.shape
        dyh = _nhwc(dy).to(xh.dtype)
        dt, st = _lib.dtype_tag(xh), stream()
        dx = torch.empty_like(xh)
        _lib.call("cgat_gate_pixels", ptr(dyh), ptr(sh), ptr(dx), dt, N, H * W, C, st)
        ds = torch.empty(N, H, W, device=xh.device, dtype=xh.dtype)
        _lib.call("cgat_chan_dot", ptr(xh), ptr(dyh), ptr(ds), dt, N * H * W, C, st)
        return dx.permute(0, 3, 1, 2), ds.view(N, 1, H, W)


def channel_gate(x, w1, b1, w2, b2):
    return _ChannelGate.apply(x, w1, b1, w2, b2)


def channel_pool(x):
    return _ChanPool.apply(x)


def pixel_gate(x, s):
    return _PixelGate.apply(x, s)
