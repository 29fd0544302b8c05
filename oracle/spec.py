"""Plain-PyTorch CPU restatement of the reference arithmetic (TEST INFRASTRUCTURE).

All functions work in whatever dtype they are given (fp32 or fp64) and use only
stock torch ops with autograd, so gradients of the oracle come from autograd.
Citations are relative to ``/root/reference``.

Canonical internal layout used by the attention functions:
``feat[n, p, node, chan]`` with ``p`` the flattened pixel index ``h*W + w``.

PINNED pieces (checked against the live reference in
``tests/test_oracle_vs_reference.py`` and through ``tests/golden``):
``adjacency_norm``, ``gat2d_layer``, ``gat1d_layer`` and the module-level wrappers
built from them, the DCGAN nets.  UNPINNED: everything about ``GATMultiHead3D``
beyond its degenerate (linear, pixel-softmax) case, and SmaAt-UNet beyond its
parameter count.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

MASK_FILL = -9e15  # pyGAT convention for masked logits (SURVEY.md F4 / lineage note)


# --------------------------------------------------------------------------------------
# a5  adjacency normalisation                         convolutional_gat/baseline_model.py:41-50,133-142
# --------------------------------------------------------------------------------------
def adjacency_norm(B: torch.Tensor) -> torch.Tensor:
    """``A_hat = D^-1/2 . minmax(B + I) . D^-1/2`` with ``D`` detached.

    baseline_model.py:44 ``adj = B + A`` (A = eye, :25), :45-47 global min/max
    normalisation, :48 ``D = Variable(diag(sum(adj, 1)), requires_grad=False)`` (so no
    gradient flows through the degree), :49 ``sqrt(inverse(D))``, :50 the two matmuls.
    """
    V = B.shape[0]
    adj = B + torch.eye(V, dtype=B.dtype, device=B.device)
    mn = adj.min()
    mx = adj.max()
    adj = (adj - mn) / (mx - mn)
    deg = adj.sum(dim=1).detach()
    r = torch.sqrt(1.0 / deg)
    return r[:, None] * adj * r[None, :]


# --------------------------------------------------------------------------------------
# a3/a4  per-pixel graph attention                      convolutional_gat/baseline_model.py:119-169
# --------------------------------------------------------------------------------------
def attention_core(
    Wh: torch.Tensor,
    a: torch.Tensor,
    A_hat: torch.Tensor,
    *,
    alpha: float = 0.2,
    softmax_axis: str = "neighbour",
    mask: Optional[torch.Tensor] = None,
    adj_transpose: bool = False,
    apply_elu: bool = True,
) -> torch.Tensor:
    """Attention on projected features ``Wh[n, p, node, co]`` -> ``out[n, p, node, co]``.

    Steps:
      2. ``e[n,p,i,j] = LeakyReLU(s1[i] + s2[j])`` with ``s1 = Wh.a[:co]``, ``s2 = Wh.a[co:]``
         -- algebraically the ``[Wh_i || Wh_j] @ a`` of baseline_model.py:128-130,162-169
         (chunks = i, alternating = j).
      3. soft-max over ``j`` ("neighbour", the 1-D layer's axis, :39) or over the pixel
         axis ``p`` ("pixel", what the 2-D layer does, :131).
      5. ``h'[i] = sum_j att[i,j] Wh[j]``                                        (:145-152)
      6. ``out[v] = ELU(sum_i h'[i] A_hat[i,v])`` (2-D layer, :154-160) or, with
         ``adj_transpose``, ``sum_i A_hat[v,i] h'[i]`` (1-D layer, :53-54).
    ``mask[i,j] == 0`` replaces the logit by ``MASK_FILL`` before the soft-max (new
    feature; all-ones reproduces the reference).
    """
    co = Wh.shape[-1]
    a = a.reshape(-1)
    s1 = Wh @ a[:co]  # [n,p,node]
    s2 = Wh @ a[co:]
    e = F.leaky_relu(s1[..., :, None] + s2[..., None, :], alpha)  # [n,p,i,j]
    if mask is not None:
        e = torch.where(mask.to(torch.bool)[None, None], e, torch.full_like(e, MASK_FILL))
    if softmax_axis == "neighbour":
        att = torch.softmax(e, dim=-1)
    elif softmax_axis == "pixel":
        att = torch.softmax(e, dim=1)
    else:
        raise ValueError(softmax_axis)
    hp = att @ Wh  # [n,p,i,co]
    M = A_hat.t() if adj_transpose else A_hat  # out[v] = sum_i hp[i] * M[i,v]
    out = torch.einsum("npic,iv->npvc", hp, M)
    return F.elu(out) if apply_elu else out


def gat2d_layer(h, W, a, B, alpha=0.2):
    """``GraphAttentionLayer2D.forward`` restated (baseline_model.py:119-160).

    ``h[N, C(=pixels), T, V]`` -> ``[N, C, T', V]``.  Pixel-axis soft-max (:131).
    """
    feat = h.permute(0, 1, 3, 2)  # [n,p,node=V,chan=T]            (:122)
    Wh = feat @ W  #                                                (:127)
    out = attention_core(Wh, a, adjacency_norm(B), alpha=alpha, softmax_axis="pixel")
    return out.permute(0, 1, 3, 2)  # [n,p,T',V]


def gat1d_layer(h, W, a, B, alpha=0.2):
    """``GraphAttentionLayer.forward`` restated (baseline_model.py:27-56).

    ``h[N, V, F]`` -> ``[N, V, F']``; soft-max over neighbours (:39); ``att <- A_hat.att``
    (:53) then ``att.Wh`` (:54).
    """
    Wh = h @ W  # [n,V,F']                                          (:35)
    out = attention_core(
        Wh[:, None], a, adjacency_norm(B), alpha=alpha, softmax_axis="neighbour", adj_transpose=True
    )
    return out[:, 0]


# --------------------------------------------------------------------------------------
# a1  builder's spec of GATMultiHead3D (UNPINNED)                         SURVEY.md appendix A.2
# --------------------------------------------------------------------------------------
def to_nodes(x: torch.Tensor, type_: str) -> torch.Tensor:
    """``x[N,H,W,T,V]`` -> ``feat[n,p,node,chan]``: spatial nodes=V chans=T; temporal nodes=T chans=V."""
    N, H, W, T, V = x.shape
    x = x.reshape(N, H * W, T, V)
    return x.permute(0, 1, 3, 2) if type_ == "spatial" else x


def from_nodes(feat: torch.Tensor, type_: str, H: int, W: int) -> torch.Tensor:
    N, P, nodes, ch = feat.shape
    if type_ == "spatial":
        return feat.permute(0, 1, 3, 2).reshape(N, H, W, ch, nodes)
    return feat.reshape(N, H, W, nodes, ch)


def node_conv3x3(feat: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], H: int, W: int):
    """Shared ``Conv2d(ci -> co, 3x3, pad 1)`` on every node's ``[ci,H,W]`` map (A.2 "conv")."""
    N, P, nodes, ci = feat.shape
    img = feat.reshape(N, H, W, nodes, ci).permute(0, 3, 4, 1, 2).reshape(N * nodes, ci, H, W)
    out = F.conv2d(img, weight, bias, padding=1)
    co = out.shape[1]
    return out.reshape(N, nodes, co, H * W).permute(0, 3, 1, 2)


def gat3d_head(
    x,
    *,
    type_,
    mapping_type,
    a,
    B,
    W=None,
    conv_weight=None,
    conv_bias=None,
    unet=None,
    alpha=0.2,
    softmax_axis="neighbour",
    mask=None,
):
    """One head of the conv-GAT layer on ``x[N,H,W,T,V]`` -> same rank (A.2)."""
    N, H, Wd, T, V = x.shape
    feat = to_nodes(x, type_)
    if mapping_type == "linear":
        Wh = feat @ W
    elif mapping_type == "conv":
        Wh = node_conv3x3(feat, conv_weight, conv_bias, H, Wd)
    elif mapping_type == "smaat_unet":
        # A.2: the shared SmaAt-UNet(ci -> co) applied to every node's [ci, H, W] map (cf. unet_model.py:25-26); the
        # nodes are folded into the batch, so train-mode BatchNorm statistics are taken over samples AND nodes
        nodes, ci = feat.shape[2], feat.shape[3]
        img = feat.reshape(N, H, Wd, nodes, ci).permute(0, 3, 4, 1, 2).reshape(N * nodes, ci, H, Wd)
        o = unet(img)
        Wh = o.reshape(N, nodes, o.shape[1], H * Wd).permute(0, 3, 1, 2)
    else:
        raise ValueError(mapping_type)
    out = attention_core(Wh, a, adjacency_norm(B), alpha=alpha, softmax_axis=softmax_axis, mask=mask)
    return from_nodes(out, type_, H, Wd)


class SpecGATHead(nn.Module):
    """Parameters of one head; names follow the in-tree layers (W, a, B) plus ``conv``."""

    def __init__(self, ci, co, n_nodes, alpha, mapping_type):
        super().__init__()
        self.alpha = alpha
        self.mapping_type = mapping_type
        if mapping_type == "linear":
            self.W = nn.Parameter(torch.empty(ci, co))
            nn.init.xavier_uniform_(self.W.data, gain=1.414)  # baseline_model.py:19-20
        elif mapping_type == "smaat_unet":
            self.unet = SpecSmaAtUNet(ci, co)
        else:
            self.conv = nn.Conv2d(ci, co, 3, padding=1)
        self.a = nn.Parameter(torch.empty(2 * co, 1))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)  # :21-22
        self.B = nn.Parameter(torch.zeros(n_nodes, n_nodes) + 1e-6)  # :24


class SpecGATStream(nn.Module):
    def __init__(self, ci, co, n_nodes, alpha, nheads, type_, mapping_type, softmax_axis, head_merge):
        super().__init__()
        self.type_, self.mapping_type = type_, mapping_type
        self.softmax_axis, self.head_merge, self.nheads = softmax_axis, head_merge, nheads
        for k in range(nheads):
            self.add_module(f"attention_{k}", SpecGATHead(ci, co, n_nodes, alpha, mapping_type))  # :191-192
        self.register_buffer("adj_mask", torch.ones(n_nodes, n_nodes, dtype=torch.uint8), persistent=False)

    def forward(self, x):
        outs = []
        for k in range(self.nheads):
            hd = getattr(self, f"attention_{k}")
            if self.mapping_type == "linear":
                kw = dict(W=hd.W)
            elif self.mapping_type == "smaat_unet":
                kw = dict(unet=hd.unet)
            else:
                kw = dict(conv_weight=hd.conv.weight, conv_bias=hd.conv.bias)
            outs.append(
                gat3d_head(
                    x, type_=self.type_, mapping_type=self.mapping_type, a=hd.a, B=hd.B, alpha=hd.alpha,
                    softmax_axis=self.softmax_axis, mask=self.adj_mask, **kw,
                )
            )
        if self.head_merge == "mean":
            return sum(outs) / self.nheads
        # concat on the channel axis: T for spatial (dim 3, cf. baseline_model.py:196), V for temporal
        return torch.cat(outs, dim=3 if self.type_ == "spatial" else 4)


class SpecGATMultiHead3D(nn.Module):
    """Spec oracle of the missing ``GAT3D.GATMultiHead3D`` (call sites convolutional_gat/model.py:21-42).

    type_ "spatial": nodes = V vertices, channels = T frames.  "temporal": nodes = T frames,
    channels = V.  "multi_stream": mean of one spatial and one temporal stream.
    """

    def __init__(self, nfeat, nhid, alpha, nheads, type_=None, mapping_type="linear", image_height=None,
                 image_width=None, n_vertices=None, softmax_axis="neighbour", head_merge="mean", **kw):
        super().__init__()
        type_ = kw.pop("type", type_)  # model.py:26 passes ``type=`` (sic)
        self.type_ = type_
        T, V = nfeat, n_vertices

        def stream(tp):
            if tp == "spatial":
                return SpecGATStream(T, nhid, V, alpha, nheads, tp, mapping_type, softmax_axis, head_merge)
            return SpecGATStream(V, V, T, alpha, nheads, tp, mapping_type, softmax_axis, head_merge)

        if type_ == "multi_stream":
            self.spatial_stream = stream("spatial")
            self.temporal_stream = stream("temporal")
        else:
            self.stream = stream(type_)

    def forward(self, x):
        if self.type_ == "multi_stream":
            return 0.5 * (self.spatial_stream(x) + self.temporal_stream(x))
        return self.stream(x)


# --------------------------------------------------------------------------------------
# a12  loss of the train step                                   convolutional_gat/train.py:131
# --------------------------------------------------------------------------------------
def train_loss(y_hat, y):
    """``MSE(y_hat, y) - 0.0005 * sum(y_hat)/numel`` (train.py:131, criterion = nn.MSELoss, :170)."""
    return F.mse_loss(y_hat, y) - 0.0005 * (y_hat.sum() / y_hat.numel())


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """One ``torch.optim.Adam(lr, weight_decay=0.01)`` update (train.py:212; L2-style decay)."""
    g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    mhat = m / (1 - beta1 ** step)
    vhat = v / (1 - beta2 ** step)
    return p - lr * mhat / (vhat.sqrt() + eps), m, v


# --------------------------------------------------------------------------------------
# a11  DCGAN nets                                                       dcgan/model.py:19-179
# --------------------------------------------------------------------------------------
def conv_same_even(x, weight, bias):
    """``Conv2d(k=4, padding="same")``: PyTorch pads total k-1=3 as left 1 / right 2 (dcgan/model.py:61-72)."""
    k = weight.shape[-1]
    total = k - 1
    lo = total // 2
    x = F.pad(x, (lo, total - lo, lo, total - lo))
    return F.conv2d(x, weight, bias)


def dcgan_generator(x, sd, prefix="", training=False, eps=1e-5):
    """``Generator.forward`` (dcgan/model.py:55-76) from a state_dict; BN in eval or batch-stat mode, dropout off."""
    for k in range(5):
        w, b = sd[f"{prefix}layers.{k}.layers.0.weight"], sd[f"{prefix}layers.{k}.layers.0.bias"]
        x = conv_same_even(x, w, b)
        if k < 4:
            p = f"{prefix}layers.{k}.layers.1."
            x = F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd[p + "weight"],
                             sd[p + "bias"], training, 0.1, eps)
            x = F.relu(x)
        else:
            x = torch.sigmoid(x)
    return x


def dcgan_frame_disc(x, sd, prefix="", training=False, eps=1e-5):
    """``FrameDiscriminator.forward`` (dcgan/model.py:171-179)."""
    x = F.leaky_relu(F.conv2d(x, sd[prefix + "conv1.weight"], None, 2, 1), 0.2)
    for k in (2, 3, 4):
        x = F.conv2d(x, sd[f"{prefix}conv{k}.weight"], None, 2, 1)
        p = f"{prefix}bn{k}."
        x = F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd[p + "weight"],
                         sd[p + "bias"], training, 0.1, eps)
        x = F.leaky_relu(x, 0.2)
    x = torch.sigmoid(F.conv2d(x, sd[prefix + "conv5.weight"], None, 1, 0))
    return x.squeeze()


def dcgan_temporal_disc(x, sd, prefix="", training=False, eps=1e-5):
    """``TemporalDiscriminator.forward`` (dcgan/model.py:79-142); last conv k4 stride 4."""
    for k in range(5):
        stride, pad = (2, 1) if k < 4 else (4, 0)
        x = F.conv2d(x, sd[f"{prefix}layers.{k}.layers.0.weight"], None, stride, pad)
        if k in (1, 2, 3):
            p = f"{prefix}layers.{k}.layers.1."
            x = F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd[p + "weight"],
                             sd[p + "bias"], training, 0.1, eps)
        x = F.leaky_relu(x, 0.2) if k < 4 else torch.sigmoid(x)
    return x.squeeze()


# --------------------------------------------------------------------------------------
# a10  SmaAt-UNet (public architecture; UNPINNED except for its parameter count)
#      call site convolutional_gat/unet_model.py:20 ``SmaAt_UNet(n_channels=4, n_classes=4)``
# --------------------------------------------------------------------------------------
class _DSConv(nn.Module):
    def __init__(self, cin, cout, kpl):
        super().__init__()
        self.depthwise = nn.Conv2d(cin, cin * kpl, 3, padding=1, groups=cin)
        self.pointwise = nn.Conv2d(cin * kpl, cout, 1)

    def forward(self, x):
        return self.pointwise(self.depthwise(x))


class _DoubleConvDS(nn.Module):
    def __init__(self, cin, cout, mid=None, kpl=1):
        super().__init__()
        mid = mid or cout
        self.double_conv = nn.Sequential(
            _DSConv(cin, mid, kpl), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
            _DSConv(mid, cout, kpl), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        )

    def forward(self, x):
        return self.double_conv(x)


class _DownDS(nn.Module):
    def __init__(self, cin, cout, kpl):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _DoubleConvDS(cin, cout, kpl=kpl))

    def forward(self, x):
        return self.maxpool_conv(x)


class _UpDS(nn.Module):
    def __init__(self, cin, cout, kpl):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = _DoubleConvDS(cin, cout, cin // 2, kpl=kpl)

    def forward(self, x1, x2):
        x1 = self.up(x1)
        dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
        x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv(torch.cat([x2, x1], dim=1))


class _ChannelAttention(nn.Module):
    def __init__(self, c, r):
        super().__init__()
        self.MLP = nn.Sequential(nn.Flatten(), nn.Linear(c, c // r), nn.ReLU(), nn.Linear(c // r, c))

    def forward(self, x):
        s = self.MLP(F.adaptive_avg_pool2d(x, 1)) + self.MLP(F.adaptive_max_pool2d(x, 1))
        return x * torch.sigmoid(s)[:, :, None, None]


class _SpatialAttention(nn.Module):
    def __init__(self, k=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, k, padding=(k - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(1)

    def forward(self, x):
        o = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
        return x * torch.sigmoid(self.bn(self.conv(o)))


class _CBAM(nn.Module):
    def __init__(self, c, r=16):
        super().__init__()
        self.channel_att = _ChannelAttention(c, r)
        self.spatial_att = _SpatialAttention(7)

    def forward(self, x):
        return self.spatial_att(self.channel_att(x))


class SpecSmaAtUNet(nn.Module):
    """Public SmaAt-UNet (kernels_per_layer=2, reduction_ratio=16, bilinear) -- 4,032,548 params at (4,4)."""

    def __init__(self, n_channels, n_classes, kernels_per_layer=2, reduction_ratio=16):
        super().__init__()
        k, r = kernels_per_layer, reduction_ratio
        self.inc = _DoubleConvDS(n_channels, 64, kpl=k)
        self.cbam1 = _CBAM(64, r)
        self.down1 = _DownDS(64, 128, k)
        self.cbam2 = _CBAM(128, r)
        self.down2 = _DownDS(128, 256, k)
        self.cbam3 = _CBAM(256, r)
        self.down3 = _DownDS(256, 512, k)
        self.cbam4 = _CBAM(512, r)
        self.down4 = _DownDS(512, 512, k)
        self.cbam5 = _CBAM(512, r)
        self.up1 = _UpDS(1024, 256, k)
        self.up2 = _UpDS(512, 128, k)
        self.up3 = _UpDS(256, 64, k)
        self.up4 = _UpDS(128, 64, k)
        self.outc = nn.Conv2d(64, n_classes, 1)

    def forward(self, x):
        x1 = self.inc(x); a1 = self.cbam1(x1)
        x2 = self.down1(x1); a2 = self.cbam2(x2)
        x3 = self.down2(x2); a3 = self.cbam3(x3)
        x4 = self.down3(x3); a4 = self.cbam4(x4)
        x5 = self.down4(x4); a5 = self.cbam5(x5)
        x = self.up1(a5, a4)
        x = self.up2(x, a3)
        x = self.up3(x, a2)
        x = self.up4(x, a1)
        return self.outc(x)


def unet_model_forward(unet: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """``UnetModel.forward`` (convolutional_gat/unet_model.py:22-29): the shared UNet per vertex, sequentially."""
    xv = x.permute(4, 0, 3, 1, 2)  # [V,B,T,H,W]                       (:24)
    acc = [unet(xv[i]) for i in range(xv.shape[0])]  #                  (:25-26)
    return torch.stack(acc).permute(1, 3, 4, 2, 0)  #                   (:27-28)


# ----------------------------------------------------------------------------------------------------------------
# KNMI loader windows (convolutional_gat/data_loaders/kmni_data_loader.py:72-127) and validation metrics
# (convolutional_gat/train.py:53-75, utils.py:135-167)
# ----------------------------------------------------------------------------------------------------------------
def kmni_windows(frames: torch.Tensor, start, *, crop=None, steps: int = 4, normalizing_max: float = 254.0,
                 power: float = 1.0):
    """``frames [L, V, H, W]`` raw integers -> ``(x, y)`` ``[N, H', W', steps, V]`` for the windows beginning at ``start``:
    ``/normalizing_max`` (:75), ``pow`` (:76), window ``i .. i+2*steps-1`` split in two (:79-94), crop (:95-96), permute
    to ``[N, H, W, T, V]`` (:121)."""
    data = torch.pow(frames.to(torch.int64) / normalizing_max, torch.tensor(power))
    xs, ys = [], []
    for i in start:
        seg = data[int(i):int(i) + 2 * steps]  # [2*steps, V, H, W]
        if crop is not None:
            seg = seg[:, :, :crop, :crop]
        xs.append(seg[:steps].permute(2, 3, 0, 1))
        ys.append(seg[steps:].permute(2, 3, 0, 1))
    return torch.stack(xs), torch.stack(ys)

# ----------------------------------------------------------------------------------------------------------------
# ARAI loader (convolutional_gat/data_loaders/arai_data_loader.py:57-93, 95-191)
# ----------------------------------------------------------------------------------------------------------------
def arai_windows(frames: torch.Tensor, start, *, downsample_size=(256, 256), steps: int = 4):
    """``frames [L, R, 1, H, W]`` floats -> ``(x, y)`` ``[N, H', W', steps, R]``: crop (:144), window ``i .. i+2*steps-1``
    split in two (:74-84), ``squeeze(3)`` + ``permute(0, 3, 4, 1, 2)`` (:89-96).  No normalisation."""
    data = frames[:, :, :, :downsample_size[0], :downsample_size[1]]
    xs, ys = [], []
    for i in start:
        chunk = data[int(i):int(i) + 2 * steps]
        xs.append(chunk[:steps].squeeze(2).permute(2, 3, 0, 1))
        ys.append(chunk[steps:].squeeze(2).permute(2, 3, 0, 1))
    return torch.stack(xs), torch.stack(ys)


def arai_batch_plan(file_lengths, batch_size: int, steps: int = 4):
    """The ``(file index, window starts)`` sequence one pass over an ARAI ``DataLoader`` yields (:98-191).  A file with
    ``L`` frames gives ``L - 2*steps + 1`` windows served ``batch_size`` at a time, the short tail included, never merged
    across files (:166-182).  Reading the LAST file raises ``should_stop_iteration`` (:156-157), so no further
    ``__get_batch`` is scheduled after it (:110-115): of the last file only the FIRST batch exists.  Whether that batch is
    handed out is a race in the reference -- ``__next__`` tests the flag (:99-101) BEFORE joining the reader thread
    (:103-104): a consumer that comes back before ``t.load`` of the last file has finished gets the batch, a slower one
    gets ``StopIteration``.  The plan here is the first outcome (what the reference produced when the golden vectors were
    generated, and the only one for a one-file folder, whose single read is synchronous, :105-107)."""
    plan = []
    n_files = len(file_lengths)
    for f, L in enumerate(file_lengths):
        nw = L - (2 * steps - 1)
        last = f == n_files - 1
        for b in range(0, nw, batch_size):
            plan.append((f, list(range(b, min(b + batch_size, nw)))))
            if last:
                break
    return plan


def val_batch_sums(y, y_hat, threshold, *, power=1.0, normalizing_max=254.0):
    """``[sum sq err, sum denormalised sq err, TP, FP, FN, #equal]`` of one batch (train.py:54-75, utils.py:135-167)."""
    y = torch.pow(y.double().float(), 1 / torch.tensor(power))
    y_hat = torch.pow(y_hat.double().float(), 1 / torch.tensor(power))
    d = (y - y_hat).double()
    yb, hb = y.clone(), y_hat.clone()
    for v in (yb, hb):  # utils.py:138-141, the two in-place assignments in their order
        v[v < threshold] = 0
        v[v >= threshold] = 1
    tp = ((hb == 1) & (yb == 1)).sum()
    fp = ((hb == 1) & (yb == 0)).sum()
    fn = ((hb == 0) & (yb == 1)).sum()
    eq = (yb == hb).sum()
    return torch.stack([(d ** 2).sum(), ((d * normalizing_max) ** 2).sum(), tp.double(), fp.double(), fn.double(),
                        eq.double()])
