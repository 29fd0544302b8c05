"""Generate tests/golden/*.pt from the LIVE, unmodified reference (TEST INFRASTRUCTURE, container only).

Run:  python -m oracle.make_golden        (needs /root/reference; writes small fp32 fixtures)

Every fixture holds seeded inputs, the reference module's state_dict, its outputs and the gradients of
``sum(out * g)`` for a seeded ``g``, so that the GPU box (which has no reference tree) can check both the
oracle restatement (oracle/spec.py) and the CUDA path against numbers the reference itself produced.
Seed 369 is the reference's only seed (dcgan/train.py:181-183); inputs are U[0,1) like the loaders emit.
"""
from __future__ import annotations

import os

import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _grads(module, out, g, inputs):
    for p in module.parameters():
        p.grad = None
    out.backward(g)
    res = {f"grad.{k}": p.grad.clone() for k, p in module.named_parameters() if p.grad is not None}
    for name, t in inputs.items():
        res[f"grad_in.{name}"] = t.grad.clone()
    return res


def gat2d_layer():
    bm = ref_loader.baseline_model()
    torch.manual_seed(369)
    N, P, T, V = 2, 400, 4, 6
    lay = bm.GraphAttentionLayer2D(T, T, V, 0.2)
    with torch.no_grad():
        lay.B.add_(torch.rand(V, V) * 0.3)  # a non-symmetric learnt adjacency
    h = torch.rand(N, P, T, V, requires_grad=True)
    with ref_loader.cpu_shim():
        out = lay(h)
    g = torch.rand_like(out) - 0.5
    fx = {"h": h.detach().clone(), "g": g, "out": out.detach().clone(), "alpha": 0.2}
    fx.update({f"sd.{k}": v.clone() for k, v in lay.state_dict().items()})
    fx.update(_grads(lay, out, g, {"h": h}))
    return fx


def gat2d_layer_init():
    """Same layer at its initial state B = 1e-6 (ties in the min/max normalisation, baseline_model.py:116)."""
    bm = ref_loader.baseline_model()
    torch.manual_seed(370)
    N, P, T, V = 2, 64, 4, 6
    lay = bm.GraphAttentionLayer2D(T, T, V, 0.2)
    h = torch.rand(N, P, T, V, requires_grad=True)
    with ref_loader.cpu_shim():
        out = lay(h)
    g = torch.rand_like(out) - 0.5
    fx = {"h": h.detach().clone(), "g": g, "out": out.detach().clone(), "alpha": 0.2}
    fx.update({f"sd.{k}": v.clone() for k, v in lay.state_dict().items()})
    fx.update(_grads(lay, out, g, {"h": h}))
    return fx


def baseline2d_model():
    bm = ref_loader.baseline_model()
    torch.manual_seed(369)
    N, H, W, T, V = 2, 20, 20, 4, 6
    model = bm.BaselineModel2D(image_width=W, image_height=H, n_vertices=V)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".B"):
                p.add_(torch.rand(V, V) * 0.2)
    x = torch.rand(N, H, W, T, V, requires_grad=True)
    with ref_loader.cpu_shim():
        out = model(x)
    g = torch.rand_like(out) - 0.5
    fx = {"x": x.detach().clone(), "g": g, "out": out.detach().clone()}
    fx.update({f"sd.{k}": v.clone() for k, v in model.state_dict().items()})
    fx.update(_grads(model, out, g, {"x": x}))
    return fx


def gat1d_layer():
    bm = ref_loader.baseline_model()
    torch.manual_seed(369)
    N, V, F_ = 3, 6, 5 * 5 * 4
    lay = bm.GraphAttentionLayer(F_, F_, V, 0.2)
    with torch.no_grad():
        lay.B.add_(torch.rand(V, V) * 0.3)
    h = torch.rand(N, V, F_, requires_grad=True)
    with ref_loader.cpu_shim():
        out = lay(h)
    g = torch.rand_like(out) - 0.5
    fx = {"h": h.detach().clone(), "g": g, "out": out.detach().clone(), "alpha": 0.2}
    fx.update({f"sd.{k}": v.clone() for k, v in lay.state_dict().items()})
    fx.update(_grads(lay, out, g, {"h": h}))
    return fx


def baseline1d_model():
    bm = ref_loader.baseline_model()
    torch.manual_seed(369)
    N, H, W, T, V = 2, 6, 6, 4, 6
    model = bm.BaselineModel(image_width=W, image_height=H, n_vertices=V)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".B"):
                p.add_(torch.rand(V, V) * 0.2)
    x = torch.rand(N, H, W, T, V, requires_grad=True)
    with ref_loader.cpu_shim():
        out = model(x)
    g = torch.rand_like(out) - 0.5
    fx = {"x": x.detach().clone(), "g": g, "out": out.detach().clone()}
    fx.update({f"sd.{k}": v.clone() for k, v in model.state_dict().items()})
    fx.update(_grads(model, out, g, {"x": x}))
    return fx


def dcgan_nets():
    dc = ref_loader.dcgan_model()
    torch.manual_seed(369)
    params = {"nc": 4, "ndf": 8}  # ndf=8 keeps the fixture small; the layer structure is that of ndf=64
    N = 3
    fx = {"params.nc": 4, "params.ndf": 8}
    x = torch.rand(N, 4, 64, 64)
    y = torch.rand(N, 4, 64, 64)
    fx["x"], fx["y"] = x, y
    for name, cls, inp in (("G", dc.Generator, x), ("FD", dc.FrameDiscriminator, y),
                           ("TD", dc.TemporalDiscriminator, torch.cat((x, y), 1))):
        net = cls(params)
        # non-trivial BN statistics/affine so that eval-mode parity means something
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.uniform_(-0.2, 0.2)
                    m.running_var.uniform_(0.5, 1.5)
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.uniform_(-0.2, 0.2)
        net.eval()  # BN running stats, Dropout2d off (dcgan/model.py:46-47)
        i = inp.clone().requires_grad_()
        out = net(i)
        g = torch.rand_like(out) - 0.5
        fx[f"{name}.out"] = out.detach().clone()
        fx[f"{name}.g"] = g
        for k, v in net.state_dict().items():
            fx[f"{name}.sd.{k}"] = v.clone()
        for k, v in _grads(net, out, g, {"inp": i}).items():
            fx[f"{name}.{k}"] = v
        # train-mode forward (batch statistics), dropout disabled by p=0 surgery on a copy is NOT done:
        # Dropout2d(0.01) is random, so train-mode parity is checked against the oracle restatement instead.
    return fx


def dcgan_step():
    """One adversarial step of dcgan/train.py:97-160 on the reference's own nets (train mode: BatchNorm batch
    statistics; Dropout2d's p set to 0 on the instantiated modules because its mask is random), with the reference's
    optimisers (Adam lr 2e-4, betas (0.5, 0.999), :195-236) and criterion (BCELoss, :224)."""
    dc = ref_loader.dcgan_model()
    torch.manual_seed(369)
    params = {"nc": 4, "ndf": 8}
    N = 4
    x = torch.rand(N, 4, 64, 64)
    y = torch.rand(N, 4, 64, 64)
    netG, netFD, netTD = dc.Generator(params), dc.FrameDiscriminator(params), dc.TemporalDiscriminator(params)
    for net in (netG, netFD, netTD):
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0
        net.train()
    fx = {"params.nc": 4, "params.ndf": 8, "x": x, "y": y}
    for name, net in (("G", netG), ("FD", netFD), ("TD", netTD)):
        for k, v in net.state_dict().items():
            fx[f"{name}.sd0.{k}"] = v.clone()
    criterion = torch.nn.BCELoss()
    opt = {n: torch.optim.Adam(net.parameters(), lr=0.0002, betas=(0.5, 0.999))
           for n, net in (("G", netG), ("FD", netFD), ("TD", netTD))}
    data = x
    # ---- dcgan/train.py:103-160 ----
    netTD.zero_grad()
    netFD.zero_grad()
    real_label = torch.zeros(N) + 1
    fake_label = torch.zeros(N)
    errFD_real = criterion(netFD(y), real_label)
    errTD_real = criterion(netTD(torch.cat((data, y), dim=1)), real_label)
    errFD_real.backward()
    errTD_real.backward()
    fake_data = netG(data)
    fake_det = fake_data.detach()
    errFD_fake = criterion(netFD(fake_det), fake_label)
    errTD_fake = criterion(netTD(torch.cat((data, fake_det), dim=1)), fake_label)
    errFD_fake.backward()
    errTD_fake.backward()
    errFD = errFD_real + errFD_fake
    errTD = errTD_real + errTD_fake
    opt["FD"].step()
    opt["TD"].step()
    netG.zero_grad()
    pred_frame = netFD(fake_data).view(-1)
    pred_temp = netTD(torch.cat((data, fake_data), dim=1)).view(-1)
    errG = criterion(pred_frame, real_label) + criterion(pred_temp, real_label)
    errG.backward()
    opt["G"].step()
    fx["errFD"], fx["errTD"], fx["errG"] = errFD.detach(), errTD.detach(), errG.detach()
    for name, net in (("G", netG), ("FD", netFD), ("TD", netTD)):
        for k, v in net.state_dict().items():
            fx[f"{name}.sd1.{k}"] = v.clone()
    return fx


def kmni_loader():
    """Batches of the reference's own KNMI ``DataLoader`` (kmni_data_loader.py:15-127) on synthetic files: raw integer
    frames ``[L, V, H, W]`` in 0..254 as the preprocessing writes them (preprocessing/kmni_dataset/__main__.py:76-110);
    sliding windows of 8 frames with stride 1, ``/254``, ``pow``, crop, permutation to ``[N, H, W, T, V]``."""
    import tempfile

    km = ref_loader.kmni_loader()
    g = torch.Generator().manual_seed(369)
    fx = {}
    with tempfile.TemporaryDirectory() as d:
        files = [torch.randint(0, 255, (L, 6, 12, 12), generator=g, dtype=torch.int64) for L in (17, 10)]
        for i, f in enumerate(files):
            torch.save(f, os.path.join(d, f"{i:010d}.pt"))
            fx[f"file{i}"] = f.to(torch.uint8)
        for tag, power in (("p1", 1.0), ("p05", 0.5)):
            dl = km.DataLoader(4, d, "cpu", crop=8, shuffle=False, power=power)
            # file 0: 17 frames -> truncated to 16 (:74) -> 9 windows: batches of 4, 4, 1; file 1: 10 -> 8 -> 1 window
            for b in range(4):
                x, y = next(dl)
                fx[f"{tag}.x{b}"], fx[f"{tag}.y{b}"] = x.contiguous().clone(), y.contiguous().clone()
    return fx


def arai_loader():
    """Every batch of the reference's own ARAI ``DataLoader`` (arai_data_loader.py:14-191) through ``get_loaders`` on a
    synthetic preprocessed folder: float frames ``[L, regions, 1, H, W]`` in files ``<k>.pt`` sorted numerically, plus
    ``metadata.json``.  Three files of 14 / 9 / 12 frames at batch size 3 (of file 2, the last, only the first batch is
    served: see ``spec.arai_batch_plan``), and a one-file folder (first batch only)."""
    import json
    import tempfile

    ar = ref_loader.arai_loader()
    g = torch.Generator().manual_seed(369)
    fx = {}
    lengths = (14, 9, 12)
    with tempfile.TemporaryDirectory() as d:
        for sub in ("training", "validation"):
            os.makedirs(os.path.join(d, sub))
        for i, L in enumerate(lengths):
            f = torch.rand(L, 5, 1, 12, 12, generator=g)
            torch.save(f, os.path.join(d, "training", f"{i}.pt"))
            fx[f"file{i}"] = f
        torch.save(fx["file0"], os.path.join(d, "validation", "0.pt"))
        with open(os.path.join(d, "metadata.json"), "w") as fh:
            json.dump({"n_regions": 5, "training": {"length": sum(lengths)}, "validation": {"length": lengths[0]}}, fh)
        train, val, _ = ar.get_loaders(3, 4, d, "cpu", downsample_size=(8, 10))
        fx["train.len"] = torch.tensor(len(train))
        for tag, dl in (("train", train), ("val", val)):
            n = 0
            for x, y in dl:
                fx[f"{tag}.x{n}"], fx[f"{tag}.y{n}"] = x.contiguous().clone(), y.contiguous().clone()
                n += 1
            fx[f"{tag}.n"] = torch.tensor(n)
    return fx


def _ref_functions(relpath, names, namespace):
    """Compile the named top-level functions of a reference file (whose module cannot be imported: missing GAT3D /
    matplotlib / torchinfo) from its UNMODIFIED source text into ``namespace``."""
    import ast

    src = open(os.path.join(ref_loader.REFERENCE_ROOT, relpath)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), relpath, "exec"), namespace)
    return namespace


class _AffineModel(torch.nn.Module):
    """Stand-in model for the validation loop: y_hat = 0.9 x + 0.02 (positive, so the 1/power root is defined)."""

    def forward(self, x):
        return 0.9 * x + 0.02


def val_metrics():
    """The reference's own ``test`` loop (convolutional_gat/train.py:28-91) and ``get_metrics`` (utils.py:135-167),
    compiled from their source, on a stand-in model and synthetic batches (one of size 1: skipped, :52)."""
    ns = _ref_functions("convolutional_gat/utils.py", ("get_metrics", "accuracy", "precision", "recall"), {"t": torch})
    ns.update({"nn": torch.nn, "tqdm": lambda it: it})
    _ref_functions("convolutional_gat/train.py", ("test",), ns)
    g = torch.Generator().manual_seed(369)
    fx = {}
    for tag, power in (("p1", 1.0), ("p05", 0.5)):
        batches = []
        for n in (3, 1, 2):
            x = torch.pow(torch.randint(0, 255, (n, 8, 8, 4, 6), generator=g).float() / 254, power)
            y = torch.pow(torch.randint(0, 40, (n, 8, 8, 4, 6), generator=g).float() / 254, power)
            batches.append((x, y))

        class Loader(list):
            pass

        loader = Loader(batches)
        loader.power = torch.tensor(power)
        loader.normalizing_max = 254
        res = ns["test"](_AffineModel(), "cpu", loader)
        for i, (x, y) in enumerate(batches):
            fx[f"{tag}.x{i}"], fx[f"{tag}.y{i}"] = x, y
        for k, v in res.items():
            fx[f"{tag}.{k}"] = torch.tensor(v, dtype=torch.float64)
        # get_metrics alone on the first batch, threshold = an arbitrary level
        acc, prec, rec = ns["get_metrics"](batches[0][1], _AffineModel()(batches[0][0]), 0.05)
        fx[f"{tag}.gm"] = torch.stack([acc.double(), prec.double(), rec.double()])
    return fx


def adjacency():
    bm = ref_loader.baseline_model()
    torch.manual_seed(369)
    fx = {}
    for V in (4, 6, 32):
        lay = bm.GraphAttentionLayer(3, 3, V, 0.2)
        with torch.no_grad():
            lay.B.add_(torch.rand(V, V) * 0.5)
        h = torch.zeros(1, V, 3)
        # run the reference lines 41-50 by calling forward and recovering A_hat from a probe: with h = 0 the
        # layer output is ELU(0) = 0, so instead recompute A_hat with the reference's own ops on its B
        B = lay.B.detach().clone().requires_grad_()
        A = torch.eye(V)
        adj = B[:, :] + A[:, :]
        adj = (adj - torch.min(adj)) / (torch.max(adj) - torch.min(adj))
        D = torch.diag(torch.sum(adj, axis=1)).detach()
        D12 = torch.sqrt(torch.inverse(D))
        ah = torch.matmul(torch.matmul(D12, adj), D12)
        g = torch.rand(V, V) - 0.5
        ah.backward(g)
        fx[f"V{V}.B"] = B.detach().clone()
        fx[f"V{V}.A_hat"] = ah.detach().clone()
        fx[f"V{V}.g"] = g
        fx[f"V{V}.grad_B"] = B.grad.clone()
    return fx


FIXTURES = {
    "gat2d_layer": gat2d_layer,
    "gat2d_layer_init": gat2d_layer_init,
    "baseline2d_model": baseline2d_model,
    "gat1d_layer": gat1d_layer,
    "baseline1d_model": baseline1d_model,
    "dcgan_nets": dcgan_nets,
    "dcgan_step": dcgan_step,
    "kmni_loader": kmni_loader,
    "arai_loader": arai_loader,
    "val_metrics": val_metrics,
    "adjacency": adjacency,
}


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not found; golden vectors can only be generated where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    for name, fn in FIXTURES.items():
        fx = fn()
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"{name:20s} {os.path.getsize(path) / 1024:8.1f} KiB  {len(fx)} entries")


if __name__ == "__main__":
    main()
