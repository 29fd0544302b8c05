"""BatchNorm2d + activation + Dropout2d on the CUDA kernels of ``csrc/norm_act_kernels.cu``.

The reference's ``ConvBlock`` (dcgan/model.py:35-52) is ``conv -> BatchNorm2d -> Dropout2d(0.01) -> activation``, the
SmaAt-UNet double convs (behind convolutional_gat/unet_model.py:20) ``conv -> BatchNorm2d -> ReLU``.  ``BatchNormAct2d``
is a drop-in for ``torch.nn.BatchNorm2d`` -- same parameters, buffers and ``state_dict`` keys (``weight``, ``bias``,
``running_mean``, ``running_var``, ``num_batches_tracked``), same train / eval semantics and running-statistics update --
that also applies the activation and the channel dropout of the block in the same pass over the tensor: two launches
forward (statistics, apply), two backward (reductions, apply), against six eager PyTorch kernels forward alone.
``ActDropout2d`` is the same without the normalisation (the blocks built with ``batchnorm=False``).
Tensors keep PyTorch's NCHW *shape* in channels_last memory = the NHWC layout the kernels (and our convs) take.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, require_cuda, stream

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """``[N, C, H, W]`` (any memory format) -> contiguous ``[N, H, W, C]`` (a view for channels_last tensors)."""
    return x.permute(0, 2, 3, 1).contiguous()


class _NormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, training, momentum, eps, act, slope, mask, sets):
        require_cuda(x)
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError(f"BatchNormAct2d serves fp32 and bf16 activations, got {x.dtype}")
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        dt = _lib.dtype_tag(xh)
        st = stream()
        norm = gamma is not None
        mean = rstd = None
        if norm:
            g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
            if training:
                if N % sets:
                    raise RuntimeError(f"BatchNormAct2d: {sets} statistic sets do not divide a batch of {N}")
                ws = torch.empty(sets * (2 * C + 2) + 2, dtype=torch.float64, device=x.device)
                mean = torch.empty(sets, C, dtype=torch.float32, device=x.device)
                rstd = torch.empty_like(mean)
                _lib.call("cgat_bn_stats_sets", ptr(xh), dt, N, H * W, C, sets, ptr(ws), ptr(mean), ptr(rstd),
                          ptr(running_mean), ptr(running_var), ptr(nbt), float(momentum), float(eps), st)
            else:
                sets = 1  # constants: one row serves every image
                mean = running_mean.detach().float().contiguous()
                rstd = (running_var.detach().float() + eps).rsqrt()
        else:
            g32 = b32 = None
            sets = 1
        y = torch.empty_like(xh)
        _lib.call("cgat_bn_act_fwd_sets", ptr(xh), ptr(y), dt, N, H * W, C, sets, ptr(mean), ptr(rstd), ptr(g32), ptr(b32),
                  ptr(mask), int(act), float(slope), st)
        ctx.cfg = (act, slope, bool(training), norm, sets)
        ctx.save_for_backward(xh, mean, rstd, g32, b32, mask)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        act, slope, training, norm, sets = ctx.cfg
        xh, mean, rstd, g32, b32, mask = ctx.saved_tensors
        N, H, W, C = xh.shape
        dyh = _nhwc(dy).to(xh.dtype)
        dx = torch.empty_like(xh)
        dg = db = ws = ss = None
        if norm:
            dg = torch.empty(C, dtype=torch.float32, device=xh.device)
            db = torch.empty_like(dg)
            ss = torch.empty(2, sets, C, dtype=torch.float32, device=xh.device)
            ws = torch.empty(sets * (2 * C + 2) + 2, dtype=torch.float64, device=xh.device)
        _lib.call("cgat_bn_act_bwd_sets", ptr(xh), ptr(dyh), ptr(dx), _lib.dtype_tag(xh), N, H * W, C, sets, ptr(mean),
                  ptr(rstd), ptr(g32), ptr(b32), ptr(mask), int(act), float(slope), int(training), ptr(ws), ptr(ss), ptr(dg),
                  ptr(db), 0, stream(), launches=2)
        return (dx.permute(0, 3, 1, 2), dg, db) + (None,) * 10


class _DropoutMask:
    """Channel-dropout masks ``[N, C]`` (0 or 1/(1-p)) from the Philox kernel; the call counter lives on the device, so a
    captured CUDA graph draws a fresh mask at every replay."""

    def __init__(self):
        self.counter = None
        self.seed = None

    def draw(self, n: int, c: int, p: float, device) -> torch.Tensor:
        if self.counter is None or self.counter.device != device:
            self.counter = torch.zeros(1, dtype=torch.int64, device=device)
            self.seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        mask = torch.empty(n, c, dtype=torch.float32, device=device)
        _lib.call("cgat_dropout2d_mask", ptr(mask), n * c, float(p), ctypes.c_uint64(self.seed), ptr(self.counter), stream())
        return mask


class BatchNormAct2d(nn.Module):
    """``torch.nn.BatchNorm2d`` (same state, same semantics) + activation + ``Dropout2d(dropout)`` in one fused op.
    Order as in the reference's ConvBlock: normalise, drop channels, activate -- for the activations used there (ReLU,
    LeakyReLU, and sigmoid only in blocks WITHOUT dropout > 0 ... see ``ActDropout2d``) dropping before or after the
    activation differs only for sigmoid, where the block order (dropout, then sigmoid) is kept by ``drop_first``."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, act=ACT_NONE, slope=0.2, dropout=0.0):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.act, self.slope, self.dropout = act, slope, dropout
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._mask = _DropoutMask()
        self.sets = 1  # statistic sets: consecutive groups of the batch normalised separately (UnetModel's vertices)

    def forward(self, x):
        if self.act == ACT_SIGMOID and self.dropout > 0 and self.training:
            raise RuntimeError("dropout before a sigmoid is served by ActDropout2d(drop_first=True)")
        mask = None
        if self.training and self.dropout > 0:
            mask = self._mask.draw(x.shape[0], x.shape[1], self.dropout, x.device)
        return _NormActFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.num_batches_tracked,
                                self.training, self.momentum, self.eps, self.act, self.slope, mask, self.sets)

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}, act={self.act}, dropout={self.dropout}"


class ActDropout2d(nn.Module):
    """``Dropout2d(dropout)`` then activation, no normalisation (ConvBlock with ``batchnorm=False``, dcgan/model.py:44-52).
    For ReLU / LeakyReLU ``act(mask * z) == mask * act(z)`` (mask >= 0), one launch; for a sigmoid the dropout is applied
    first as its own pass (``sigmoid(0) = 0.5`` for a dropped channel, exactly what the reference computes)."""

    def __init__(self, act=ACT_NONE, slope=0.2, dropout=0.0):
        super().__init__()
        self.act, self.slope, self.dropout = act, slope, dropout
        self._mask = _DropoutMask()

    def forward(self, x):
        mask = None
        if self.training and self.dropout > 0:
            mask = self._mask.draw(x.shape[0], x.shape[1], self.dropout, x.device)
        none = (None,) * 5
        if mask is not None and self.act == ACT_SIGMOID:
            x = _NormActFn.apply(x, *none, False, 0.0, 0.0, ACT_NONE, 0.0, mask, 1)
            mask = None
        if mask is None and self.act == ACT_NONE:
            return x
        return _NormActFn.apply(x, *none, False, 0.0, 0.0, self.act, self.slope, mask, 1)

    def extra_repr(self):
        return f"act={self.act}, dropout={self.dropout}"


def set_dropout(module: nn.Module, p: float):
    """Set the channel-dropout probability of every fused block under ``module`` (tests pin parity at p = 0)."""
    for m in module.modules():
        if isinstance(m, (BatchNormAct2d, ActDropout2d)):
            m.dropout = p
