"""Stock-PyTorch (ATen / cuDNN / cuBLAS or CPU) restatements of the reference's modules, used ONLY as the timed
baseline rows of ``bench.py`` and ``tools/bench_config5.py`` (TEST / BASELINE INFRASTRUCTURE -- the product never
imports this).

The reference tree itself does not travel to the GPU box, so each row is rebuilt here from the oracle's functions
(``oracle/spec.py``, every one of them pinned to the live reference by ``tests/test_oracle_vs_reference.py`` /
``tests/golden``) -- i.e. exactly the ATen calls a user of the reference would run on that device:

* R1 ``BaselineModel2D``  (convolutional_gat/baseline_model.py:200-233)  two single-head 2-D GAT layers + tanh
* R2 ``BaselineModel``    (:236-270)                                       two single-head 1-D GAT layers + tanh
* R3 conv-GAT model of BASELINE config 2  (``spec.SpecGATMultiHead3D``; the upstream layer is missing)
* R4 DCGAN adversarial step (dcgan/train.py:97-160 on dcgan/model.py:19-179), nn.Conv2d / BatchNorm2d / Dropout2d
"""
from __future__ import annotations

import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import spec


class _GATParams(nn.Module):
    """W, a, B of one reference layer with the reference initialisation (baseline_model.py:19-25, :111-117)."""

    def __init__(self, fin, fout, n_vertices):
        super().__init__()
        self.W = nn.Parameter(torch.empty(fin, fout))
        self.a = nn.Parameter(torch.empty(2 * fout, 1))
        nn.init.xavier_uniform_(self.W, gain=1.414)
        nn.init.xavier_uniform_(self.a, gain=1.414)
        self.B = nn.Parameter(torch.full((n_vertices, n_vertices), 1e-6))


class Baseline2D(nn.Module):
    def __init__(self, n_vertices=6, time_steps=4):
        super().__init__()
        self.l1, self.l2 = _GATParams(time_steps, time_steps, n_vertices), _GATParams(time_steps, time_steps, n_vertices)

    def forward(self, x):
        b, h, w, t, v = x.shape
        z = x.reshape(b, h * w, t, v)
        for l in (self.l1, self.l2):
            z = spec.gat2d_layer(z, l.W, l.a, l.B)
        return torch.tanh(z.view(b, h, w, t, v))


class Baseline1D(nn.Module):
    def __init__(self, image=20, n_vertices=6, time_steps=4):
        super().__init__()
        f = time_steps * image * image
        self.l1, self.l2 = _GATParams(f, f, n_vertices), _GATParams(f, f, n_vertices)

    def forward(self, x):
        b, h, w, t, v = x.shape
        z = x.reshape(b, h * w * t, v).permute(0, 2, 1)
        for l in (self.l1, self.l2):
            z = spec.gat1d_layer(z, l.W, l.a, l.B)
        return torch.tanh(z.reshape(b, h, w, t, v))


def _block(cin, cout, k, stride=1, padding=0, bias=True, bn=True, act="relu"):
    layers = [nn.Conv2d(cin, cout, k, stride, padding, bias=bias)]
    if bn:
        layers.append(nn.BatchNorm2d(cout))
    layers.append(nn.Dropout2d(0.01))
    layers.append({"relu": nn.ReLU(), "lrelu": nn.LeakyReLU(0.2), "sigmoid": nn.Sigmoid()}[act])
    return nn.Sequential(*layers)


def dcgan_nets(nc=4, ndf=64):
    """Generator / FrameDiscriminator / TemporalDiscriminator as stock torch modules (dcgan/model.py:55-179)."""
    G = nn.Sequential(_block(nc, 8 * nc, 4, padding="same"), _block(8 * nc, 4 * nc, 4, padding="same"),
                      _block(4 * nc, 2 * nc, 4, padding="same"), _block(2 * nc, nc, 4, padding="same"),
                      _block(nc, nc, 4, padding="same", bn=False, act="sigmoid"))

    def disc(cin, last_stride, dropout):
        chans = [cin, ndf, 2 * ndf, 4 * ndf, 8 * ndf]
        layers = []
        for i in range(4):
            b = _block(chans[i], chans[i + 1], 4, 2, 1, bias=False, bn=i > 0, act="lrelu")
            if not dropout:
                b = nn.Sequential(*[m for m in b if not isinstance(m, nn.Dropout2d)])
            layers.append(b)
        last = [nn.Conv2d(8 * ndf, 1, 4, last_stride, 0, bias=False)] + ([nn.Dropout2d(0.01)] if dropout else [])
        layers.append(nn.Sequential(*last, nn.Sigmoid()))
        return nn.Sequential(*layers)

    return G, disc(nc, 1, False), disc(2 * nc, 4, True)


def dcgan_step(G, FD, TD, oG, oFD, oTD, x, y):
    """One batch of dcgan/train.py:97-160: D(real) + D(fake) -> step both discriminators; G through both -> step G."""
    bce = nn.BCELoss()

    def crit(pred, label):  # BCELoss refuses to run under autocast: the criterion itself stays fp32 (dcgan/train.py:224)
        with torch.autocast(pred.device.type, enabled=False):
            return bce(pred.float(), label)

    n = x.shape[0]
    real, fake = torch.ones(n, device=x.device), torch.zeros(n, device=x.device)
    TD.zero_grad()
    FD.zero_grad()
    crit(FD(y).reshape(-1).float(), real).backward()
    crit(TD(torch.cat((x, y), 1)).reshape(-1).float(), real).backward()
    g = G(x)
    gd = g.detach()
    crit(FD(gd).reshape(-1).float(), fake).backward()
    crit(TD(torch.cat((x, gd), 1)).reshape(-1).float(), fake).backward()
    oFD.step()
    oTD.step()
    G.zero_grad()
    err = crit(FD(g).reshape(-1).float(), real) + crit(TD(torch.cat((x, g), 1)).reshape(-1).float(), real)
    err.backward()
    oG.step()
    return err.detach()


def time_fn(fn, device, warmup=1, reps=3):
    """Seconds per call: CUDA events on a CUDA device, best-of wall clock on the CPU."""
    for _ in range(warmup):
        fn()
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / reps
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def fwd_bwd(model, x):
    def run():
        for p in model.parameters():
            p.grad = None
        model(x).sum().backward()
    return run


def reference_rows(device, batch_small=4, batch_gpu=64, autocast=False, seed=369):
    """samples/s of rows R1-R4 on ``device`` (R3 at ``batch_gpu`` on CUDA, ``batch_small`` on the CPU)."""
    dev = torch.device(device)
    cuda = dev.type == "cuda"
    torch.manual_seed(seed)
    rows = {}
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if (cuda and autocast) else (lambda: torch.autocast("cpu", enabled=False))
    reps, warm = (10, 3) if cuda else (3, 1)

    def timed(fn, n):
        with ctx():
            sec = time_fn(fn, dev, warm, reps)
        return {"samples_per_s": n / sec, "ms_per_step": sec * 1e3, "batch": n}

    n = batch_small
    x20 = torch.rand(n, 20, 20, 4, 6, device=dev)
    rows["R1_BaselineModel2D_20x20_fwd_bwd"] = timed(fwd_bwd(Baseline2D().to(dev), x20), n)
    rows["R2_BaselineModel_20x20_fwd_bwd"] = timed(fwd_bwd(Baseline1D().to(dev), x20), n)
    nb = batch_gpu if cuda else batch_small
    m = spec.SpecGATMultiHead3D(4, 4, 0.2, 3, type_="temporal", mapping_type="conv", n_vertices=6).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0.01)
    x, y = torch.rand(nb, 64, 64, 4, 6, device=dev), torch.rand(nb, 64, 64, 4, 6, device=dev)

    def step3():
        opt.zero_grad()
        spec.train_loss(m(x), y).backward()
        opt.step()

    rows["R3_convgat_config2_train_step"] = timed(step3, nb)
    nd = batch_gpu if cuda else batch_small
    G, FD, TD = (net.to(dev) for net in dcgan_nets())
    mk = lambda net: torch.optim.Adam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    oG, oFD, oTD = mk(G), mk(FD), mk(TD)
    xd, yd = torch.rand(nd, 4, 64, 64, device=dev), torch.rand(nd, 4, 64, 64, device=dev)
    rows["R4_dcgan_adversarial_step"] = timed(lambda: dcgan_step(G, FD, TD, oG, oFD, oTD, xd, yd), nd)
    if cuda:
        # BASELINE config 4 (stress): UnetModel [2,128,128,4,8] -- one shared SmaAt-UNet per vertex in a Python loop
        # (convolutional_gat/unet_model.py:22-29), full train step (train.py:129-133, :212), stock torch modules
        unet = spec.SpecSmaAtUNet(4, 4).to(dev)
        ou = torch.optim.Adam(unet.parameters(), lr=1e-3, weight_decay=0.01)
        xu, yu = torch.rand(2, 128, 128, 4, 8, device=dev), torch.rand(2, 128, 128, 4, 8, device=dev)

        def step_u():
            ou.zero_grad()
            spec.train_loss(spec.unet_model_forward(unet, xu), yu).backward()
            ou.step()

        rows["config4_unet_model_train_step"] = timed(step_u, 2)
    return rows
