"""autograd Functions over the C ABI: fused graph attention, NHWC convolution, loss, Adam.

Each Function's forward/backward is one or two kernel launches on torch's current stream; nothing
here synchronises, so a whole train step can be captured in a CUDA graph.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import AttnDesc, ConvDesc, check, dtype_tag, lib, ptr, require_cuda, stream


# ----------------------------------------------------------------------------------------------
# fused graph attention (K4 / K5)
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class AttnConfig:
    nodes: int
    ci: int
    co: int
    heads: int
    layout: int  # _lib.LAYOUT_*
    proj: int  # _lib.PROJ_*
    merge: int  # _lib.MERGE_*
    pix_per_sample: int
    alpha: float = 0.2
    softmax_axis: str = "neighbour"  # or "pixel" (baseline_model.py:131 compat)
    adj_transpose: bool = False  # True: the 1-D layer's  A_hat . att  (baseline_model.py:53)
    apply_elu: bool = True

    def desc(self, n_pix: int, dtype: int) -> AttnDesc:
        return AttnDesc(n_pix, self.pix_per_sample, self.nodes, self.ci, self.co, self.heads, self.layout, self.proj,
                        self.merge, dtype, int(self.apply_elu), self.alpha)

    @property
    def in_rec(self) -> int:
        return self.heads * self.nodes * self.co if self.proj == _lib.PROJ_PRE else self.nodes * self.ci

    @property
    def out_rec(self) -> int:
        return self.nodes * self.co if self.merge == _lib.MERGE_MEAN else self.heads * self.nodes * self.co


class _GraphAttention(torch.autograd.Function):
    """``out = attention(inp; W, a, B)`` for all heads of one stream in a single launch.

    inp  [n_pix, in_rec]   fp32/bf16, contiguous pixel records
    W    [heads, ci, co]   fp32 or None (PROJ_PRE)
    a    [heads, 2*co]     fp32
    B    [heads, nodes, nodes] fp32 (raw learnable adjacency, baseline_model.py:116)
    mask [nodes, nodes]    uint8 or None
    """

    @staticmethod
    def forward(ctx, inp, W, a, B, mask, cfg: AttnConfig):
        require_cuda(inp, W, a, B, mask)
        L = lib()
        inp = inp.contiguous()
        n_pix = inp.shape[0]
        if inp.shape[1] != cfg.in_rec:
            raise RuntimeError(f"attention input record has {inp.shape[1]} elements, expected {cfg.in_rec}")
        dt = dtype_tag(inp)
        Wc = None if W is None else W.detach().float().contiguous()
        ac = a.detach().float().contiguous()
        Bc = B.detach().float().contiguous()
        mc = None if mask is None else mask.to(torch.uint8).contiguous()
        st = stream()
        adj = torch.empty_like(Bc)
        _lib.call("cgat_adj_norm_fwd", ptr(Bc), ptr(adj), cfg.heads, cfg.nodes, int(cfg.adj_transpose), st)
        d = cfg.desc(n_pix, dt)
        stats = None
        if cfg.softmax_axis == "pixel":
            n_samples = n_pix // cfg.pix_per_sample
            stats = torch.empty(n_samples, cfg.heads, 2, cfg.nodes * cfg.nodes, device=inp.device, dtype=torch.float32)
            _lib.call("cgat_attn_pixstats", ctypes.byref(d), ptr(inp), ptr(Wc), ptr(ac), ptr(mc), ptr(stats), st)
        out = torch.empty(n_pix, cfg.out_rec, device=inp.device, dtype=inp.dtype)
        _lib.call("cgat_attn_fwd", ctypes.byref(d), ptr(inp), ptr(out), ptr(Wc), ptr(ac), ptr(adj), ptr(mc), ptr(stats), st)
        ctx.cfg = cfg
        ctx.has_W = W is not None
        ctx.save_for_backward(inp, Wc, ac, Bc, adj, mc, stats)
        return out

    @staticmethod
    def backward(ctx, dout):
        cfg: AttnConfig = ctx.cfg
        inp, Wc, ac, Bc, adj, mc, stats = ctx.saved_tensors
        L = lib()
        dout = dout.contiguous()
        n_pix = inp.shape[0]
        d = cfg.desc(n_pix, dtype_tag(inp))
        st = stream()
        # one zeroed fp32 buffer for all parameter-gradient accumulators
        nW = cfg.heads * cfg.ci * cfg.co if ctx.has_W else 0
        na = cfg.heads * 2 * cfg.co
        nadj = cfg.heads * cfg.nodes * cfg.nodes
        nb = (n_pix // cfg.pix_per_sample) * nadj if stats is not None else 0
        acc = torch.zeros(nW + na + nadj + nb, device=inp.device, dtype=torch.float32)
        gW = acc[:nW].view(cfg.heads, cfg.ci, cfg.co) if ctx.has_W else None
        ga = acc[nW:nW + na].view(cfg.heads, 2 * cfg.co)
        gadj = acc[nW + na:nW + na + nadj].view(cfg.heads, cfg.nodes, cfg.nodes)
        bstats = acc[nW + na + nadj:] if stats is not None else None
        if stats is not None:
            _lib.call("cgat_attn_pixstats_bwd", ctypes.byref(d), ptr(inp), ptr(dout), ptr(Wc), ptr(ac), ptr(adj), ptr(mc),
                                           ptr(stats), ptr(bstats), st)
        din = torch.empty_like(inp)
        _lib.call("cgat_attn_bwd", ctypes.byref(d), ptr(inp), ptr(dout), ptr(din), ptr(Wc), ptr(ac), ptr(adj), ptr(mc),
                              ptr(stats), ptr(bstats), ptr(gW), ptr(ga), ptr(gadj), st)
        gB = torch.empty_like(Bc)
        _lib.call("cgat_adj_norm_bwd", ptr(Bc), ptr(gadj), ptr(gB), cfg.heads, cfg.nodes, int(cfg.adj_transpose), st)
        return din, gW, ga, gB, None, None


def graph_attention(inp, W, a, B, mask, cfg: AttnConfig):
    return _GraphAttention.apply(inp, W, a, B, mask, cfg)


def adjacency_norm(B: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """Normalised adjacency of baseline_model.py:41-50 for ``B[heads, nodes, nodes]`` (no autograd)."""
    require_cuda(B)
    Bc = B.detach().float().contiguous()
    out = torch.empty_like(Bc)
    _lib.call("cgat_adj_norm_fwd", ptr(Bc), ptr(out), Bc.shape[0], Bc.shape[1], int(transpose), stream())
    return out


# ----------------------------------------------------------------------------------------------
# NHWC convolution (K1 / K2 / K3)
# ----------------------------------------------------------------------------------------------
IMPL_DIRECT, IMPL_TC, IMPL_AUTO = 0, 1, -1


def _conv_desc(n, h, w, cin, cout, kh, kw, stride, pad_top, pad_left, ho, wo, dtype, act=0, groups=1) -> ConvDesc:
    return ConvDesc(n, h, w, cin, cout, kh, kw, stride, pad_top, pad_left, ho, wo, dtype, act, groups)


def _pick(d: ConvDesc, which: int, impl: int) -> int:
    if impl != IMPL_AUTO:
        return impl
    return IMPL_TC if lib().cgat_conv_tc_supported(ctypes.byref(d), which) else IMPL_DIRECT


def _workspace(d: ConvDesc, which: int, impl: int, device):
    if impl != IMPL_TC:
        return None
    nbytes = lib().cgat_conv_workspace_bytes(ctypes.byref(d), which)
    return torch.empty(max(16, nbytes), dtype=torch.uint8, device=device) if nbytes else None


class _Conv2dNHWC(torch.autograd.Function):
    """``y[n,ho,wo,cout] = act(conv(x[n,h,w,cin], w[cout,kh,kw,cin]) + bias)``.

    ``pad = (top, left, bottom, right)``; activations are fused in the epilogue, their derivative is
    applied to ``dy`` here before dgrad/wgrad.
    """

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, act, impl, groups=1):
        require_cuda(x, w, bias)
        x = x.contiguous()
        n, h, wd, cin = x.shape
        cout, kh, kw, cin2 = w.shape
        if cin != cin2 * groups or cout % groups:
            raise RuntimeError(f"conv: input has {cin} channels, weight expects {cin2} x {groups} groups")
        pt, pl, pb, pr = pad
        ho = (h + pt + pb - kh) // stride + 1
        wo = (wd + pl + pr - kw) // stride + 1
        dt = dtype_tag(x)
        wk = w.detach().to(x.dtype).contiguous()
        bk = None if bias is None else bias.detach().float().contiguous()
        d = _conv_desc(n, h, wd, cin, cout, kh, kw, stride, pt, pl, ho, wo, dt, act, groups)
        y = torch.empty(n, ho, wo, cout, device=x.device, dtype=x.dtype)
        im = _pick(d, 0, impl)
        ws = _workspace(d, 0, im, x.device)
        _lib.call("cgat_conv2d_fprop", ctypes.byref(d), ptr(x), ptr(wk), ptr(bk), ptr(y), im, ptr(ws), stream(),
                  launches=2 if im == IMPL_TC else 1)
        ctx.d = d
        ctx.impl = impl
        ctx.has_bias = bias is not None
        ctx.w_dtype = w.dtype
        ctx.save_for_backward(x, wk, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wk, y = ctx.saved_tensors
        d: ConvDesc = ctx.d
        if d.act == _lib.ACT_RELU:
            dy = dy * (y > 0)
        elif d.act == _lib.ACT_LRELU:
            dy = dy * torch.where(y > 0, 1.0, 0.2).to(dy.dtype)
        elif d.act == _lib.ACT_SIGMOID:
            yf = y.float()
            dy = (dy.float() * yf * (1 - yf)).to(dy.dtype)
        dy = dy.contiguous()
        st = stream()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            im = _pick(d, 1, ctx.impl)
            ws = _workspace(d, 1, im, x.device)
            _lib.call("cgat_conv2d_dgrad", ctypes.byref(d), ptr(dy), ptr(wk), ptr(dx), im, ptr(ws), st,
                      launches=2 if im == IMPL_TC else 1)
        dw = torch.empty(wk.shape, device=x.device, dtype=torch.float32)
        db = torch.empty(d.cout, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        im = _pick(d, 2, ctx.impl)
        ws = _workspace(d, 2, im, x.device)
        _lib.call("cgat_conv2d_wgrad", ctypes.byref(d), ptr(x), ptr(dy), ptr(dw), ptr(db), im, ptr(ws), st,
                  launches=2 if db is not None else 1)
        return dx, dw.to(ctx.w_dtype), db, None, None, None, None, None


def conv2d_nhwc(x, w_krsc, bias=None, stride=1, pad=(0, 0, 0, 0), act=0, impl=IMPL_AUTO, groups=1):
    """NHWC convolution with KRSC weights ``[cout, kh, kw, cin/groups]`` through the CUDA kernels (no cuDNN)."""
    return _Conv2dNHWC.apply(x, w_krsc, bias, stride, tuple(pad), act, impl, groups)


# ----------------------------------------------------------------------------------------------
# train-step pieces (convolutional_gat/train.py:131, :212)
# ----------------------------------------------------------------------------------------------
def loss_and_grad(y_hat: torch.Tensor, y: torch.Tensor, lam: float = 0.0005, grad_scale: float = 1.0,
                  loss_out: Optional[torch.Tensor] = None, dy_out: Optional[torch.Tensor] = None):
    """``MSE(y_hat,y) - lam*mean(y_hat)`` and its gradient w.r.t. ``y_hat`` in one launch.

    Returns ``(loss[1] fp32 (accumulated into ``loss_out`` if given), d loss / d y_hat)``.
    """
    require_cuda(y_hat, y)
    y_hat = y_hat.contiguous()
    y = y.contiguous().to(y_hat.dtype)
    if loss_out is None:
        loss_out = torch.zeros(1, device=y_hat.device, dtype=torch.float32)
    if dy_out is None:
        dy_out = torch.empty_like(y_hat)
    _lib.call("cgat_loss_fwd_bwd", ptr(y_hat), ptr(y), ptr(dy_out), ptr(loss_out), y_hat.numel(), lam, grad_scale,
                                  dtype_tag(y_hat), stream())
    return loss_out, dy_out


def adam_step_(param, grad, m, v, step_dev, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, grad_scale=1.0):
    """In-place ``torch.optim.Adam`` update of flat fp32 buffers; ``step_dev`` is an int64 device scalar."""
    require_cuda(param, grad, m, v, step_dev)
    _lib.call("cgat_adam_step", ptr(param), ptr(grad), ptr(m), ptr(v), ptr(step_dev), param.numel(), lr, beta1, beta2,
                               eps, weight_decay, grad_scale, stream())
