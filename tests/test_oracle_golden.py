"""The oracle restatement (oracle/spec.py) against golden vectors produced by the LIVE reference
(oracle/make_golden.py).  CPU only; this is what pins the oracle."""
import pytest
import torch

from oracle import spec
from util import close, golden, sd_of


def test_adjacency_norm_matches_reference():
    fx = golden("adjacency")
    for V in (4, 6, 32):
        B = fx[f"V{V}.B"].clone().requires_grad_()
        ah = spec.adjacency_norm(B)
        close(ah, fx[f"V{V}.A_hat"], msg=f"A_hat V={V}")
        ah.backward(fx[f"V{V}.g"])
        close(B.grad, fx[f"V{V}.grad_B"], rtol=1e-4, atol=1e-5, msg=f"grad_B V={V}")


@pytest.mark.parametrize("name", ["gat2d_layer", "gat2d_layer_init"])
def test_gat2d_layer_matches_reference(name):
    fx = golden(name)
    sd = sd_of(fx)
    h = fx["h"].clone().requires_grad_()
    W, a, B = (sd[k].clone().requires_grad_() for k in ("W", "a", "B"))
    out = spec.gat2d_layer(h, W, a, B, fx["alpha"])
    close(out, fx["out"], msg="out")
    out.backward(fx["g"])
    close(h.grad, fx["grad_in.h"], atol=1e-5, msg="dh")
    close(W.grad, fx["grad.W"], atol=1e-4, msg="dW")
    close(a.grad, fx["grad.a"], atol=1e-4, msg="da")
    close(B.grad, fx["grad.B"], atol=1e-4, msg="dB")


def test_gat1d_layer_matches_reference():
    fx = golden("gat1d_layer")
    sd = sd_of(fx)
    h = fx["h"].clone().requires_grad_()
    W, a, B = (sd[k].clone().requires_grad_() for k in ("W", "a", "B"))
    out = spec.gat1d_layer(h, W, a, B, fx["alpha"])
    close(out, fx["out"], msg="out")
    out.backward(fx["g"])
    close(h.grad, fx["grad_in.h"], atol=1e-5, msg="dh")
    close(W.grad, fx["grad.W"], atol=1e-4, msg="dW")
    close(a.grad, fx["grad.a"], atol=1e-4, msg="da")
    close(B.grad, fx["grad.B"], atol=1e-4, msg="dB")


def test_baseline2d_model_matches_reference():
    fx = golden("baseline2d_model")
    sd = sd_of(fx)
    x = fx["x"]
    N, H, W, T, V = x.shape
    h = x.reshape(N, H * W, T, V)
    for layer in ("hidden_layer", "output_layer"):
        p = f"{layer}.attention_0."
        h = spec.gat2d_layer(h, sd[p + "W"], sd[p + "a"], sd[p + "B"])
    close(torch.tanh(h.view(N, H, W, T, V)), fx["out"], msg="BaselineModel2D out")


def test_baseline1d_model_matches_reference():
    fx = golden("baseline1d_model")
    sd = sd_of(fx)
    x = fx["x"]
    N, H, W, T, V = x.shape
    h = x.reshape(N, H * W * T, V).permute(0, 2, 1)  # baseline_model.py:266
    for layer in ("hidden_layer", "output_layer"):
        p = f"{layer}.attention_0."
        h = spec.gat1d_layer(h, sd[p + "W"], sd[p + "a"], sd[p + "B"])
    close(torch.tanh(h.reshape(N, H, W, T, V)), fx["out"], msg="BaselineModel out")  # raw view, :269


def test_dcgan_nets_match_reference():
    fx = golden("dcgan_nets")
    x, y = fx["x"], fx["y"]
    for name, fn, inp in (("G", spec.dcgan_generator, x), ("FD", spec.dcgan_frame_disc, y),
                          ("TD", spec.dcgan_temporal_disc, torch.cat((x, y), 1))):
        sd = sd_of(fx, f"{name}.sd.")
        i = inp.clone().requires_grad_()
        sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v)
               for k, v in sd.items()}
        out = fn(i, sdg)
        close(out, fx[f"{name}.out"], msg=f"{name} out")
        out.backward(fx[f"{name}.g"])
        close(i.grad, fx[f"{name}.grad_in.inp"], atol=1e-6, msg=f"{name} dx")
        for k, v in sdg.items():
            if v.requires_grad:
                close(v.grad, fx[f"{name}.grad.{k}"], atol=1e-5, msg=f"{name} d{k}")


def test_smaat_unet_parameter_count():
    """compare_models/results/results.json:18 -- the only pin the reference holds for SmaAt-UNet."""
    net = spec.SpecSmaAtUNet(4, 4)
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 4_032_548
    out = spec.unet_model_forward(net.eval(), torch.rand(1, 16, 16, 4, 2))
    assert out.shape == (1, 16, 16, 4, 2)


def test_spec_gat3d_degenerates_to_pinned_2d_layer():
    """A.2 parity ladder (i): linear mapping + pixel soft-max + all-ones mask + concat == GraphAttentionLayer2D."""
    fx = golden("gat2d_layer")
    sd = sd_of(fx)
    h = fx["h"]  # [N,P,T,V]
    N, P, T, V = h.shape
    layer = spec.SpecGATMultiHead3D(T, T, 0.2, 1, type_="spatial", mapping_type="linear", n_vertices=V,
                                    softmax_axis="pixel", head_merge="concat")
    layer.stream.attention_0.load_state_dict(sd)
    out = layer(h.reshape(N, 20, 20, T, V))
    close(out.reshape(N, P, T, V), fx["out"], msg="spec 3D vs reference 2D")


@pytest.mark.parametrize("tag,power", [("p1", 1.0), ("p05", 0.5)])
def test_kmni_windows_match_reference_loader(tag, power):
    """oracle/spec.kmni_windows against batches of the reference's own DataLoader (shuffle off)."""
    fx = golden("kmni_loader")
    # file 0: 17 frames -> 16 -> 9 windows in batches of 4, 4, 1; file 1: 10 -> 8 -> 1 window
    plan = [("file0", range(0, 4)), ("file0", range(4, 8)), ("file0", range(8, 9)), ("file1", range(0, 1))]
    for b, (f, starts) in enumerate(plan):
        x, y = spec.kmni_windows(fx[f], list(starts), crop=8, power=power)
        assert torch.equal(x, fx[f"{tag}.x{b}"]) and torch.equal(y, fx[f"{tag}.y{b}"])


def test_arai_windows_and_batch_plan_match_reference_loader():
    """oracle/spec.arai_windows + arai_batch_plan against every batch of the reference's own ARAI DataLoader."""
    fx = golden("arai_loader")
    plan = spec.arai_batch_plan([14, 9, 12], 3)
    assert len(plan) == int(fx["train.n"]) == 5
    for b, (f, starts) in enumerate(plan):
        x, y = spec.arai_windows(fx[f"file{f}"], starts, downsample_size=(8, 10))
        assert torch.equal(x, fx[f"train.x{b}"]) and torch.equal(y, fx[f"train.y{b}"])
    plan = spec.arai_batch_plan([14], 4)  # one-file folder: its first batch only
    assert len(plan) == int(fx["val.n"]) == 1
    x, y = spec.arai_windows(fx["file0"], plan[0][1], downsample_size=(8, 10))
    assert torch.equal(x, fx["val.x0"]) and torch.equal(y, fx["val.y0"])


@pytest.mark.parametrize("tag,power", [("p1", 1.0), ("p05", 0.5)])
def test_val_metrics_match_reference_test_loop(tag, power):
    """oracle/spec.val_batch_sums, folded as train.py:76-91 does, against the reference's own ``test`` / ``get_metrics``."""
    fx = golden("val_metrics")
    tot = torch.zeros(5, dtype=torch.float64)
    n_seen = 0
    for i in range(3):
        x, y = fx[f"{tag}.x{i}"], fx[f"{tag}.y{i}"]
        if len(x) <= 1:
            continue
        y_hat = 0.9 * x + 0.02
        uniq = torch.unique(torch.pow(y, 1 / torch.tensor(power)))
        thr = uniq[int(len(uniq) * 0.5)]
        s = spec.val_batch_sums(y, y_hat, thr, power=power)
        n, per = len(x), y[0].numel()
        tot += torch.stack([s[0] / per, s[5] / per, s[2] / (s[2] + s[3]) * n, s[2] / (s[2] + s[4]) * n, s[1] / per])
        n_seen += n
    for k, v in zip(("val_loss", "val_acc", "val_prec", "val_rec", "val_denorm_mse"), tot / n_seen):
        close(v, fx[f"{tag}.{k}"], rtol=2e-5, atol=1e-7, msg=k)
    s = spec.val_batch_sums(fx[f"{tag}.y0"], 0.9 * fx[f"{tag}.x0"] + 0.02, 0.05)
    n, per = 3, fx[f"{tag}.y0"][0].numel()
    close(torch.stack([s[5] / per, s[2] / (s[2] + s[3]) * n, s[2] / (s[2] + s[4]) * n]), fx[f"{tag}.gm"], rtol=1e-6,
          atol=1e-7, msg="get_metrics")


def test_dcgan_step_fixture_is_consistent():
    """The adversarial-step fixture holds the reference's state before and after its step; the nets moved."""
    fx = golden("dcgan_step")
    for name in ("G", "FD", "TD"):
        sd0, sd1 = sd_of(fx, f"{name}.sd0."), sd_of(fx, f"{name}.sd1.")
        assert sd0.keys() == sd1.keys()
        assert any((sd0[k].float() - sd1[k].float()).abs().max() > 0 for k in sd0)
    assert all(torch.isfinite(fx[k]) for k in ("errFD", "errTD", "errG"))
