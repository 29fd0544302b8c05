"""nn.Modules of the conv-GAT hot path on the CUDA kernels.

Class names, constructor kwargs, attribute names and state_dict keys mirror the reference so that
the modules drop into ``convolutional_gat/train.py`` unchanged:

* ``GraphAttentionLayer2D`` / ``GATMultiHead2D`` / ``GraphAttentionLayer`` / ``GATMultiHead`` --
  the in-tree layers of ``convolutional_gat/baseline_model.py:13-197`` (state_dict: ``W``, ``a``, ``B``;
  heads registered as ``attention_{i}``).
* ``GATMultiHead3D`` -- the layer ``convolutional_gat/model.py:21-42`` imports from the missing
  ``GAT3D`` sub-module; semantics defined in DESIGN.md (spec: oracle/spec.py, PARITY UNPINNED).

All arithmetic runs in libcgat_b200.so; there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .functional import (AttnConfig, adjacency_norm_autograd, conv2d_nhwc, gat1d_core, gat_stream, graph_attention,
                         IMPL_AUTO)


def _xavier(shape):
    p = nn.Parameter(torch.empty(*shape))
    nn.init.xavier_uniform_(p.data, gain=1.414)  # baseline_model.py:19-22
    return p


class _HeadParams(nn.Module):
    """Parameters of one attention head (keys ``W``/``conv.*``, ``a``, ``B`` as in baseline_model.py:111-116)."""

    def __init__(self, ci: int, co: int, n_nodes: int, alpha: float, mapping_type: str = "linear"):
        super().__init__()
        self.in_features, self.out_features, self.alpha = ci, co, alpha
        self.mapping_type = mapping_type
        if mapping_type == "linear":
            self.W = _xavier((ci, co))
        elif mapping_type == "conv":
            self.conv = nn.Conv2d(ci, co, 3, padding=1)  # parameter container only; the math runs in our kernels
        elif mapping_type == "smaat_unet":
            # the shared SmaAt-UNet(ci -> co) per node (cf. unet_model.py:20-26); imported here: that module imports us
            from convolutional_gat.GAT3D.smaat_unet.SmaAt_UNet import SmaAt_UNet

            self.unet = SmaAt_UNet(n_channels=ci, n_classes=co)
        else:
            raise ValueError(f"mapping_type {mapping_type!r} not supported here")
        self.a = _xavier((2 * co, 1))
        self.B = nn.Parameter(torch.zeros(n_nodes, n_nodes) + 1e-6)  # baseline_model.py:24

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


def _conv_stream_served(N, H, W, nodes, ci, co, heads, layout, merge, alpha) -> bool:
    """True when the one-launch stream op serves this conv-mapped shape: either the fused layer kernels
    (cgat_layer_fwd/bwd) or the tcgen05 fprop + wgrad pair on the dense block-diagonal conv."""
    import ctypes

    from . import functional as F

    L = _lib.lib()
    if F.FUSED_LAYER:
        ld = _lib.LayerDesc(N, H, W, nodes, ci, co, heads, layout, merge, 1, alpha)
        if L.cgat_layer_supported(ctypes.byref(ld)):
            return True
    cin, cout = nodes * ci, heads * nodes * co
    if cin % 8 or cout % 8 or cout > 128:
        return False
    cd = _lib.ConvDesc(N, H, W, cin, cout, 3, 3, 1, 1, 1, H, W, _lib.BF16, 0, 1)
    # exactly what _GATStreamFn calls: the packed (resident-weight) fprop, wgrad and -- for stacked layers -- dgrad
    return bool(L.cgat_conv_stream_supported(ctypes.byref(cd), 1))


class _GATStream(nn.Module):
    """All heads of one stream, one fused kernel launch per direction."""

    def __init__(self, ci, co, n_nodes, alpha, nheads, type_, mapping_type="linear", softmax_axis="neighbour",
                 head_merge="mean", conv_impl=IMPL_AUTO):
        super().__init__()
        if type_ not in ("spatial", "temporal"):
            raise ValueError(type_)
        if head_merge not in ("mean", "concat"):
            raise ValueError(head_merge)
        if softmax_axis not in ("neighbour", "pixel"):
            raise ValueError(softmax_axis)
        self.ci, self.co, self.n_nodes, self.alpha, self.nheads = ci, co, n_nodes, alpha, nheads
        self.type_, self.mapping_type = type_, mapping_type
        self.softmax_axis, self.head_merge, self.conv_impl = softmax_axis, head_merge, conv_impl
        self.attentions = [_HeadParams(ci, co, n_nodes, alpha, mapping_type) for _ in range(nheads)]
        for i, att in enumerate(self.attentions):
            self.add_module(f"attention_{i}", att)  # baseline_model.py:191-192
        # all-ones mask == the reference's dense attention; not part of the state_dict
        self.register_buffer("adj_mask", torch.ones(n_nodes, n_nodes, dtype=torch.uint8), persistent=False)

    # -- dense block-diagonal expansion of the shared per-node conv (see DESIGN.md "node conv") --
    def _dense_conv_params(self, other: int):
        w = torch.stack([h.conv.weight for h in self.attentions])  # [k, co, ci, 3, 3]
        b = torch.stack([h.conv.bias for h in self.attentions])  # [k, co]
        k, co, ci = w.shape[:3]
        eye = torch.eye(other, device=w.device, dtype=w.dtype)
        if self.type_ == "spatial":
            # rows (k,u,v)  cols (kh,kw,(t,v'))   value w[k,u,t,kh,kw] d(v,v')
            dense = torch.einsum("kuthw,vx->kuvhwtx", w, eye).reshape(k * co * other, 3, 3, ci * other)
            bias = b[:, :, None].expand(k, co, other).reshape(-1)
        else:
            # rows (k,t,u)  cols (kh,kw,(t',v))   value w[k,u,v,kh,kw] d(t,t')
            dense = torch.einsum("kuvhw,tx->ktuhwxv", w, eye).reshape(k * other * co, 3, 3, other * ci)
            bias = b[:, None, :].expand(k, other, co).reshape(-1)
        return dense, bias

    def _mask_arg(self):
        """The neighbour mask for the kernels, or None while it is all ones (the reference's dense attention: the
        kernels then skip the mask arithmetic).  Cached on the buffer's version counter (one host sync per change)."""
        m = self.adj_mask
        key = (m.data_ptr(), m._version)
        if getattr(self, "_mask_key", None) != key:
            self._mask_all_ones = bool(m.all().item())
            self._mask_key = key
        return None if self._mask_all_ones else m

    def _train_cfg(self, x):
        N, H, W, T, V = x.shape
        spatial = self.type_ == "spatial"
        nodes = V if spatial else T
        return AttnConfig(nodes=nodes, ci=self.ci, co=self.co, heads=self.nheads,
                          layout=_lib.LAYOUT_SPATIAL if spatial else _lib.LAYOUT_TEMPORAL, proj=_lib.PROJ_PRE,
                          merge=_lib.MERGE_MEAN if self.head_merge == "mean" else _lib.MERGE_CONCAT,
                          pix_per_sample=H * W, alpha=self.alpha, softmax_axis=self.softmax_axis)

    def train_step_supported(self, x: torch.Tensor) -> bool:
        """Whether ``fused_train_step`` serves this stream for inputs like ``x`` (see functional.gat_stream_train)."""
        from .functional import layer_train_supported

        if x.dim() != 5 or not all(p.dtype == torch.float32 for p in self.parameters()):
            return False
        return layer_train_supported(x, self._train_cfg(x), self.mapping_type)

    def fused_train_step(self, x, y, lam, loss_out, mse_out=None, acc=None, x_planar=None, scratch=None, precision="fp16x2",
                         adam=None, clear=None, mirror=None):
        """forward + ``MSE - lam*mean`` loss + backward of a model that is just this stream (train.py:130-132):
        accumulates the loss into ``loss_out`` and the gradients into the parameters' ``.grad`` buffers."""
        from .functional import gat_stream_train

        params = []
        for h in self.attentions:
            params += [h.conv.weight, h.conv.bias, h.a, h.B]
        gat_stream_train(x, y, self._train_cfg(x), self._mask_arg(), params, lam, loss_out, mse_out, acc, x_planar=x_planar,
                         scratch=scratch, precision=precision, adam=adam, clear=clear, mirror=mirror)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x[N,H,W,T,V]`` -> ``[N,H,W,T,V]`` (mean merge) or heads concatenated on the channel axis."""
        if x.dim() != 5:
            raise RuntimeError(f"expected x[N,H,W,T,V], got shape {tuple(x.shape)}")
        N, H, W, T, V = x.shape
        spatial = self.type_ == "spatial"
        nodes, ch = (V, T) if spatial else (T, V)
        if nodes != self.n_nodes or ch != self.ci:
            raise RuntimeError(
                f"{self.type_} stream built for nodes={self.n_nodes}, channels={self.ci}; input has nodes={nodes}, "
                f"channels={ch}")
        layout = _lib.LAYOUT_SPATIAL if spatial else _lib.LAYOUT_TEMPORAL
        merge = _lib.MERGE_MEAN if self.head_merge == "mean" else _lib.MERGE_CONCAT
        conv = self.mapping_type in ("conv", "smaat_unet")
        fused = (self.softmax_axis == "neighbour" and self.mapping_type != "smaat_unet"
                 and all(p.dtype == torch.float32 for p in self.parameters()))
        if fused and conv:
            fused = (x.dtype == torch.bfloat16 and self.conv_impl == IMPL_AUTO
                     and _conv_stream_served(N, H, W, nodes, self.ci, self.co, self.nheads, layout, merge, self.alpha))
        if fused:
            # one prepare launch + kernels + one param-grad launch (functional._GATStreamFn)
            cfg = AttnConfig(nodes=nodes, ci=self.ci, co=self.co, heads=self.nheads, layout=layout,
                             proj=_lib.PROJ_PRE if conv else _lib.PROJ_LINEAR, merge=merge, pix_per_sample=H * W,
                             alpha=self.alpha, softmax_axis="neighbour")
            params = []
            for h in self.attentions:
                params += [h.conv.weight, h.conv.bias, h.a, h.B] if conv else [h.W, h.a, h.B]
            out = gat_stream(x, cfg, self.mapping_type, self._mask_arg(), params)
        else:
            a = torch.stack([h.a.reshape(-1) for h in self.attentions])
            B = torch.stack([h.B for h in self.attentions])
            if not conv:
                Wt = torch.stack([h.W for h in self.attentions])
                inp = x.reshape(N * H * W, T * V)
                proj = _lib.PROJ_LINEAR
            elif self.mapping_type == "smaat_unet":
                # nodes folded into the batch: [N*nodes, ci, H, W] through the head's UNet (its convs run in K1-K3)
                xn = (x.permute(0, 4, 3, 1, 2) if spatial else x.permute(0, 3, 4, 1, 2)).reshape(N * nodes, self.ci, H, W)
                per_head = []
                for h in self.attentions:
                    yv = h.unet(xn).reshape(N, nodes, self.co, H, W)
                    per_head.append(yv.permute(0, 3, 4, 2, 1) if spatial else yv.permute(0, 3, 4, 1, 2))
                inp = torch.stack(per_head, dim=3).reshape(N * H * W, -1).contiguous()
                Wt = None
                proj = _lib.PROJ_PRE
            elif nodes > 8:
                # many nodes (BASELINE config 4: V = 32 / 64): the block-diagonal dense conv would be 1/nodes dense, so
                # the shared conv runs once over the nodes folded into the batch, [N*nodes, H, W, ci] -> [.., co]
                xn = (x.permute(0, 4, 1, 2, 3) if spatial else x.permute(0, 3, 1, 2, 4)).reshape(N * nodes, H, W, self.ci)
                per_head = []
                for h in self.attentions:
                    y = conv2d_nhwc(xn.contiguous(), h.conv.weight.permute(0, 2, 3, 1), h.conv.bias, stride=1,
                                    pad=(1, 1, 1, 1), impl=self.conv_impl).view(N, nodes, H, W, self.co)
                    # head sub-record: (node, c) at c*nodes + node (spatial) or node*co + c (temporal)
                    per_head.append(y.permute(0, 2, 3, 4, 1) if spatial else y.permute(0, 2, 3, 1, 4))
                inp = torch.stack(per_head, dim=3).reshape(N * H * W, -1)
                Wt = None
                proj = _lib.PROJ_PRE
            else:
                dense, bias = self._dense_conv_params(nodes)
                wh = conv2d_nhwc(x.reshape(N, H, W, T * V), dense, bias, stride=1, pad=(1, 1, 1, 1), impl=self.conv_impl)
                inp = wh.reshape(N * H * W, -1)
                Wt = None
                proj = _lib.PROJ_PRE
            cfg = AttnConfig(nodes=nodes, ci=self.ci, co=self.co, heads=self.nheads, layout=layout, proj=proj,
                             merge=merge, pix_per_sample=H * W, alpha=self.alpha, softmax_axis=self.softmax_axis)
            out = graph_attention(inp, Wt, a, B, self.adj_mask, cfg)
        hm = 1 if self.head_merge == "mean" else self.nheads
        if spatial:
            return out.view(N, H, W, hm * self.co, V)
        return out.view(N, H, W, T, hm * self.co)


class GATMultiHead3D(nn.Module):
    """The conv graph-attention layer of ``convolutional_gat/model.py:21-42`` (source missing upstream).

    ``type_`` "spatial": nodes = V vertices, channels = T frames (``nfeat -> nhid``).
    "temporal": nodes = T frames, channels = V (``n_vertices -> n_vertices``).
    "multi_stream": mean of a spatial and a temporal stream.
    ``mapping_type`` "linear" (``Wh = X.W``, baseline_model.py:127), "conv" (shared 3x3 conv per node) or
    "smaat_unet" (shared SmaAt-UNet(ci -> co) per node, the nodes folded into the batch; cf. unet_model.py:20-26).
    Extra keyword-only knobs (not in the reference call sites): ``softmax_axis`` ("neighbour" | "pixel"),
    ``head_merge`` ("mean" keeps the output shape equal to the input shape, which train.py:131 needs;
    "concat" follows baseline_model.py:196).  The legacy kwarg ``type=`` (model.py:26) is accepted.
    """

    def __init__(self, nfeat, nhid, alpha, nheads, type_=None, mapping_type="linear", image_height=None,
                 image_width=None, n_vertices=None, *, softmax_axis="neighbour", head_merge="mean",
                 conv_impl=IMPL_AUTO, **kwargs):
        super().__init__()
        if "type" in kwargs:
            type_ = kwargs.pop("type")
        if kwargs:
            raise TypeError(f"unexpected kwargs {sorted(kwargs)}")
        if type_ not in ("spatial", "temporal", "multi_stream"):
            raise ValueError(f"type_ must be spatial|temporal|multi_stream, got {type_!r}")
        self.type_, self.mapping_type = type_, mapping_type
        self.image_height, self.image_width, self.n_vertices = image_height, image_width, n_vertices
        T, V = nfeat, n_vertices

        def stream(tp):
            if tp == "spatial":
                return _GATStream(T, nhid, V, alpha, nheads, tp, mapping_type, softmax_axis, head_merge, conv_impl)
            return _GATStream(V, V, T, alpha, nheads, tp, mapping_type, softmax_axis, head_merge, conv_impl)

        if type_ == "multi_stream":
            self.spatial_stream = stream("spatial")
            self.temporal_stream = stream("temporal")
        else:
            self.stream = stream(type_)

    def forward(self, x):
        if self.type_ == "multi_stream":
            return 0.5 * (self.spatial_stream(x) + self.temporal_stream(x))
        return self.stream(x)


# ------------------------------------------------------------------------------------------------
# drop-ins for the in-tree layers (convolutional_gat/baseline_model.py)
# ------------------------------------------------------------------------------------------------
class GraphAttentionLayer2D(_HeadParams):
    """``baseline_model.GraphAttentionLayer2D`` (:105-179): per-pixel attention, soft-max over the PIXEL axis."""

    def __init__(self, in_features, out_features, n_vertices, alpha):
        super().__init__(in_features, out_features, n_vertices, alpha, "linear")

    def forward(self, h):
        return _gat2d_forward([self], h)


def _gat2d_forward(heads, h):
    if h.dim() != 4:
        raise RuntimeError(f"expected h[N,C,T,V], got {tuple(h.shape)}")
    N, C, T, V = h.shape
    h0 = heads[0]
    cfg = AttnConfig(nodes=V, ci=T, co=h0.out_features, heads=len(heads), layout=_lib.LAYOUT_SPATIAL,
                     proj=_lib.PROJ_LINEAR, merge=_lib.MERGE_CONCAT, pix_per_sample=C, alpha=h0.alpha,
                     softmax_axis="pixel")
    Wt = torch.stack([m.W for m in heads])
    a = torch.stack([m.a.reshape(-1) for m in heads])
    B = torch.stack([m.B for m in heads])
    out = graph_attention(h.reshape(N * C, T * V), Wt, a, B, None, cfg)
    return out.view(N, C, len(heads) * h0.out_features, V)  # cat on dim=2 (:196)


class GATMultiHead2D(nn.Module):
    """``baseline_model.GATMultiHead2D`` (:182-197); all heads in one launch."""

    def __init__(self, nfeat, nhid, n_vertices, alpha, nheads):
        super().__init__()
        self.attentions = [GraphAttentionLayer2D(nfeat, nhid, n_vertices, alpha) for _ in range(nheads)]
        for i, att in enumerate(self.attentions):
            self.add_module(f"attention_{i}", att)

    def forward(self, x):
        return _gat2d_forward(self.attentions, x)


class GraphAttentionLayer(_HeadParams):
    """``baseline_model.GraphAttentionLayer`` (:13-75): features flattened per vertex, soft-max over neighbours.

    ``Wh = h.W`` runs in the conv kernels as a 1x1 convolution over the ``[N, V]`` grid of rows; everything after it
    (scores, soft-max, ``A_hat . attention`` (:53), aggregation, ELU) in ``cgat_gat1d_fwd/bwd``.  fp32.
    """

    def __init__(self, in_features, out_features, n_vertices, alpha):
        super().__init__(in_features, out_features, n_vertices, alpha, "linear")

    def forward(self, h):
        if h.dim() == 4:  # :28-30
            N, C, T, V = h.shape
            h = h.permute(0, 3, 1, 2).contiguous().view(N, V, C * T)
        N, V, F_ = h.shape
        x = h.float().reshape(1, N, V, F_)
        w_krsc = self.W.t().reshape(self.out_features, 1, 1, F_)
        Wh = conv2d_nhwc(x, w_krsc).reshape(N, V, self.out_features)  # :35
        adj = adjacency_norm_autograd(self.B[None])[0]  # :41-50
        return gat1d_core(Wh, self.a, adj, None, self.alpha)


class GATMultiHead(nn.Module):
    """``baseline_model.GATMultiHead`` (:78-102): heads concatenated on the feature axis."""

    def __init__(self, nfeat, nhid, n_vertices, alpha, nheads):
        super().__init__()
        self.attentions = [GraphAttentionLayer(nfeat, nhid, n_vertices, alpha) for _ in range(nheads)]
        for i, att in enumerate(self.attentions):
            self.add_module(f"attention_{i}", att)

    def forward(self, x):
        return torch.cat([att(x) for att in self.attentions], dim=-1)  # :93
