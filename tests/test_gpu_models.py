"""GPU parity of the conv stacks around the GAT layer: DCGAN nets vs golden vectors of the live reference,
SmaAt-UNet / UnetModel vs the oracle restatement."""
import pytest
import torch

from oracle import spec
from util import close, close_frac, golden, sd_of

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", ["G", "FD", "TD"])
def test_dcgan_net_vs_reference_golden_fp32(name):
    from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator

    fx = golden("dcgan_nets")
    params = {"nc": fx["params.nc"], "ndf": fx["params.ndf"]}
    cls = {"G": Generator, "FD": FrameDiscriminator, "TD": TemporalDiscriminator}[name]
    net = cls(params)
    net.load_state_dict(sd_of(fx, f"{name}.sd."))  # the reference's own state_dict keys
    net = net.to(DEV).eval()
    x, y = fx["x"], fx["y"]
    inp = {"G": x, "FD": y, "TD": torch.cat((x, y), 1)}[name].to(DEV).requires_grad_()
    out = net(inp)
    close(out, fx[f"{name}.out"], rtol=1e-4, atol=1e-5, msg=f"{name} out")
    out.backward(fx[f"{name}.g"].to(DEV))
    close(inp.grad, fx[f"{name}.grad_in.inp"], rtol=1e-4, atol=1e-6, msg=f"{name} dx")
    for k, p in net.named_parameters():
        ref = fx[f"{name}.grad.{k}"]
        close(p.grad, ref, rtol=1e-4, atol=1e-5 * max(1.0, ref.abs().max().item()), msg=f"{name} d{k}")


def test_dcgan_generator_bf16_tensor_core_path():
    """bf16 activations: layers with cin % 8 == 0 run on tcgen05 (IMPL_AUTO), the 4-channel ones on the direct kernel."""
    from dcgan.model import Generator

    fx = golden("dcgan_nets")
    net = Generator({"nc": 4, "ndf": 8})
    net.load_state_dict(sd_of(fx, "G.sd."))
    net = net.to(DEV).eval()
    x = fx["x"].bfloat16()
    ref = spec.dcgan_generator(x.float(), sd_of(fx, "G.sd."))
    out = net(x.to(DEV))
    assert out.dtype == torch.bfloat16
    close(out, ref, rtol=2e-2, atol=2e-2, msg="G bf16")


def test_dcgan_train_mode_batchnorm_matches_oracle():
    """Train-mode (batch statistics) forward with dropout disabled: parity with the oracle restatement."""
    from dcgan.model import FrameDiscriminator

    fx = golden("dcgan_nets")
    net = FrameDiscriminator({"nc": 4, "ndf": 8})
    sd = sd_of(fx, "FD.sd.")
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    y = fx["y"]
    ref = spec.dcgan_frame_disc(y, sd, training=True)
    close(net(y.to(DEV)), ref, rtol=1e-4, atol=1e-5, msg="FD train-mode")


def test_smaat_unet_matches_oracle_fp32():
    from convolutional_gat.GAT3D.smaat_unet.SmaAt_UNet import SmaAt_UNet

    torch.manual_seed(3)
    ref = spec.SpecSmaAtUNet(4, 4).eval()
    ours = SmaAt_UNet(4, 4)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).eval()
    assert sum(p.numel() for p in ours.parameters()) == 4_032_548  # compare_models/results/results.json:18
    x = torch.rand(2, 4, 32, 32)
    with torch.no_grad():
        close(ours(x.to(DEV)), ref(x), rtol=1e-3, atol=1e-4, msg="SmaAt-UNet out")


def test_unet_model_per_vertex_loop_and_grads():
    from convolutional_gat.unet_model import UnetModel

    torch.manual_seed(4)
    ref = spec.SpecSmaAtUNet(4, 4).train()
    ours = UnetModel(image_width=16, image_height=16, n_vertices=3, attention_type="unet")
    ours.unet.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).train()
    x = torch.rand(2, 16, 16, 4, 3)
    # train mode: per-vertex BatchNorm statistics and V sequential running-stat updates (unet_model.py:25-26)
    with torch.no_grad():
        out_r = spec.unet_model_forward(ref, x)
        out_o = ours(x.to(DEV))
    close(out_o, out_r, rtol=5e-3, atol=5e-4, msg="UnetModel out (train mode)")  # 40+ fp32 layers
    close(ours.unet.inc.double_conv[1].running_mean, ref.inc.double_conv[1].running_mean, rtol=1e-4, atol=1e-6,
          msg="running_mean after V updates")
    # gradients are compared in eval mode: at this size the bottleneck BatchNorm normalises over 2 values in
    # train mode, which amplifies fp32 rounding by 1/sigma and makes the comparison ill-conditioned
    ref.eval()
    ours.eval()
    xr = x.clone().requires_grad_()
    out_r = spec.unet_model_forward(ref, xr)
    g = torch.rand_like(out_r) - 0.5
    out_r.backward(g)
    xo = x.to(DEV).requires_grad_()
    out_o = ours(xo)
    close(out_o, out_r.detach(), rtol=5e-3, atol=5e-4, msg="UnetModel out (eval)")
    out_o.backward(g.to(DEV))
    close(xo.grad, xr.grad, rtol=1e-2, atol=1e-4 * max(1.0, xr.grad.abs().max().item()), msg="UnetModel dx")
    pr = dict(ref.named_parameters())
    for k, p in ours.unet.named_parameters():
        close(p.grad, pr[k].grad, rtol=1e-2, atol=1e-3 * max(1.0, pr[k].grad.abs().max().item()), msg=f"d{k}")


def test_unet_model_batched_vertices_equal_the_per_vertex_loop():
    """UnetModel runs its V per-vertex passes (unet_model.py:25-26) as one batched pass with per-vertex BatchNorm statistic
    sets.  Train-mode gradients of this 40-layer fp32 net carry ~0.5 % of rounding noise (ReLU / arg-max decisions next
    to a tie flip with the last bit of a batch mean), so the literal loop and the batched pass are both measured against
    the fp64 oracle: the batched pass must be as close to it as the loop is, and the BatchNorm buffers must agree."""
    import copy

    from convolutional_gat.unet_model import UnetModel

    torch.manual_seed(5)
    a = UnetModel(image_width=64, image_height=64, n_vertices=3, attention_type="unet").to(DEV).train()
    ref = spec.SpecSmaAtUNet(4, 4)
    ref.load_state_dict({k: v.cpu() for k, v in a.unet.state_dict().items()})
    ref = ref.double().train()
    b = copy.deepcopy(a)
    b.batched = False
    x = torch.rand(2, 64, 64, 4, 3, device=DEV)
    g = torch.rand(2, 64, 64, 4, 3, device=DEV) - 0.5
    xr = x.cpu().double().requires_grad_()
    out_r = spec.unet_model_forward(ref, xr)
    out_r.backward(g.cpu().double())
    res = []
    for m in (a, b):
        xi = x.clone().requires_grad_()
        out = m(xi)
        out.backward(g)
        res.append((out.detach(), xi.grad))
    torch.cuda.synchronize()

    def err(u, v):
        return ((u.cpu().double() - v).abs().max() / v.abs().max().clamp_min(1e-30)).item()

    e_out = [err(r[0], out_r.detach()) for r in res]
    e_dx = [err(r[1], xr.grad) for r in res]
    assert e_out[0] < 1e-4 and e_out[0] <= 2 * e_out[1] + 1e-6, e_out
    assert e_dx[0] < 2e-2 and e_dx[0] <= 2 * e_dx[1] + 1e-5, e_dx
    pr, pb = dict(ref.named_parameters()), dict(b.unet.named_parameters())
    scale = max(p.grad.abs().max().item() for p in pr.values())
    tot_a = tot_b = 0.0
    for k, p in a.unet.named_parameters():
        gr = pr[k].grad
        if gr.abs().max().item() < 1e-6 * scale:  # biases in front of a BatchNorm: the true gradient is zero
            continue
        # absolute noise of a cancelling sum (e.g. the single dgamma of a CBAM spatial BatchNorm2d(1): +-0.05 in either
        # path, whatever the value) is bounded against the model's gradient scale; the paths are compared in aggregate
        da = (p.grad.cpu().double() - gr).abs().max().item()
        db = (pb[k].grad.cpu().double() - gr).abs().max().item()
        assert da <= 3e-2 * gr.abs().max().item() + 5e-3 * scale, (k, da, db, gr.abs().max().item(), scale)
        tot_a += da / (gr.abs().max().item() + 1e-2 * scale)
        tot_b += db / (gr.abs().max().item() + 1e-2 * scale)
    assert tot_a <= 1.5 * tot_b + 1e-3, (tot_a, tot_b)
    sb, sr = dict(b.named_buffers()), dict(ref.named_buffers())
    for k, v in a.named_buffers():
        close(v.float(), sb[k].float(), rtol=1e-4, atol=1e-6, msg=k)
        close(v.float(), sr[k[len("unet."):]].float(), rtol=1e-4, atol=1e-6, msg="oracle " + k)
    assert int(a.unet.inc.double_conv[1].num_batches_tracked) == 3


def test_model_registry_and_train_signature():
    """train.py:198-205: model_classes[model_type](image_width=, image_height=, n_vertices=, attention_type=, mapping_type=)."""
    from convolutional_gat.utils import model_classes

    for mt in ("temporal", "spatial", "multi_stream"):
        m = model_classes[mt](image_width=16, image_height=16, n_vertices=6, attention_type=mt, mapping_type="conv").to(DEV)
        assert m.mapping_type == "conv"
        x = torch.rand(2, 16, 16, 4, 6, device=DEV)
        y_hat = m(x)
        assert y_hat.shape == x.shape  # train.py:131 needs MSE(y_hat, y)
        sd = m.state_dict()
        m.load_state_dict(sd)


def test_dcgan_adversarial_step_vs_reference_golden():
    """BASELINE config 5: one adversarial step (dcgan/train.py:97-160) on our nets vs the live reference's step
    (tests/golden/dcgan_step.pt: same initial state_dicts, batch, optimisers; Dropout2d p = 0)."""
    from cgat.norm_act import set_dropout
    from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator
    from dcgan.train import adversarial_step, default_criterion, make_optimizers

    fx = golden("dcgan_step")
    params = {"nc": fx["params.nc"], "ndf": fx["params.ndf"]}
    nets = {"G": Generator(params), "FD": FrameDiscriminator(params), "TD": TemporalDiscriminator(params)}
    for name, net in nets.items():
        net.load_state_dict(sd_of(fx, f"{name}.sd0."))
        set_dropout(net, 0.0)
        net.to(DEV).train()
    oG, oFD, oTD = make_optimizers(nets["G"], nets["FD"], nets["TD"])
    errFD, errTD, errG, _ = adversarial_step(netG=nets["G"], netFD=nets["FD"], netTD=nets["TD"], optimizerG=oG,
                                             optimizerFD=oFD, optimizerTD=oTD, criterion=default_criterion(),
                                             x=fx["x"].to(DEV), y=fx["y"].to(DEV))
    close(errFD, fx["errFD"], rtol=1e-4, atol=1e-5, msg="errFD")
    close(errTD, fx["errTD"], rtol=1e-4, atol=1e-5, msg="errTD")
    close(errG, fx["errG"], rtol=1e-4, atol=1e-5, msg="errG")
    # Adam's first step moves every parameter by ~lr * sign(grad): compare the updated state (parameters and
    # BatchNorm running statistics) -- an entry whose gradient is ~0 may flip its sign, hence the 2*lr slack
    for name, net in nets.items():
        sd1 = sd_of(fx, f"{name}.sd1.")
        for k, v in net.state_dict().items():
            close(v, sd1[k], rtol=1e-4, atol=4.1e-4 if v.dtype.is_floating_point and "running" not in k else 1e-5,
                  msg=f"{name}.{k} after the step")
        moved = sum(int((v.float().cpu() - fx[f"{name}.sd0.{k}"].float()).abs().max() > 0) for k, v in net.state_dict().items())
        assert moved > 0


def test_dcgan_graphed_step_vs_reference_golden():
    """The CUDA-graph replay of the adversarial step (dcgan.train.GraphedAdversarialStep) is the same first step from the
    same state as the live reference's (tests/golden/dcgan_step.pt): construction's warm-up steps leave no trace."""
    from cgat.norm_act import set_dropout
    from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator
    from dcgan.train import GraphedAdversarialStep, default_criterion, make_optimizers

    fx = golden("dcgan_step")
    params = {"nc": fx["params.nc"], "ndf": fx["params.ndf"]}
    nets = {"G": Generator(params), "FD": FrameDiscriminator(params), "TD": TemporalDiscriminator(params)}
    for name, net in nets.items():
        net.load_state_dict(sd_of(fx, f"{name}.sd0."))
        set_dropout(net, 0.0)
        net.to(DEV).train()
    oG, oFD, oTD = make_optimizers(nets["G"], nets["FD"], nets["TD"], capturable=True)
    x, y = fx["x"].to(DEV), fx["y"].to(DEV)
    step = GraphedAdversarialStep(netG=nets["G"], netFD=nets["FD"], netTD=nets["TD"], optimizerG=oG, optimizerFD=oFD,
                                  optimizerTD=oTD, criterion=default_criterion(), x=x, y=y)
    for name, net in nets.items():  # construction restored the caller's state
        for k, v in net.state_dict().items():
            assert torch.equal(v.cpu(), fx[f"{name}.sd0.{k}"]), f"{name}.{k} changed by construction"
    errFD, errTD, errG, _ = step(x, y)
    close(errFD, fx["errFD"], rtol=1e-4, atol=1e-5, msg="errFD")
    close(errTD, fx["errTD"], rtol=1e-4, atol=1e-5, msg="errTD")
    close(errG, fx["errG"], rtol=1e-4, atol=1e-5, msg="errG")
    for name, net in nets.items():
        sd1 = sd_of(fx, f"{name}.sd1.")
        for k, v in net.state_dict().items():
            close(v, sd1[k], rtol=1e-4, atol=4.1e-4 if v.dtype.is_floating_point and "running" not in k else 1e-5,
                  msg=f"{name}.{k} after the replayed step")


@pytest.mark.parametrize("name", ["FD", "TD"])
def test_dcgan_discriminators_bf16_stride2_on_tensor_cores(name):
    """bf16: the k=4 stride-2 convs run as space-to-depth + 2x2 stride-1 convs on the tcgen05 kernels (cgat.conv_layers).
    Forward against the fp32 golden of the live reference at the bf16 bar; forward and every gradient against the SAME
    bf16 net on the direct CUDA-core kernels (identical products, different summation order), which isolates the
    regrouping from the bf16 rounding that five stacked layers accumulate."""
    from cgat import _lib
    from cgat.conv_layers import Conv2d
    from cgat.functional import IMPL_AUTO, IMPL_DIRECT
    from dcgan.model import FrameDiscriminator, TemporalDiscriminator
    import cgat.functional as F

    fx = golden("dcgan_nets")
    params = {"nc": fx["params.nc"], "ndf": fx["params.ndf"]}
    x, y = fx["x"], fx["y"]
    res = {}
    for mode in ("tc", "direct"):
        net = {"FD": FrameDiscriminator, "TD": TemporalDiscriminator}[name](params)
        net.load_state_dict(sd_of(fx, f"{name}.sd."))
        net = net.to(DEV).eval()
        for m in net.modules():
            if isinstance(m, Conv2d):
                m.impl = IMPL_AUTO if mode == "tc" else IMPL_DIRECT
        inp = {"FD": y, "TD": torch.cat((x, y), 1)}[name].to(DEV).bfloat16().requires_grad_()
        names, orig = [], _lib.call

        def spy(n, *a, **k):
            names.append((n, a[5] if n == "cgat_conv2d_fprop" else None))
            return orig(n, *a, **k)

        F._lib.call = spy
        try:
            out = net(inp)
        finally:
            F._lib.call = orig
        n_tc = sum(1 for n, impl in names if n == "cgat_conv2d_fprop" and impl == 1)
        assert (n_tc >= 3) if mode == "tc" else (n_tc == 0), names
        out.float().backward(fx[f"{name}.g"].to(DEV))
        res[mode] = dict({k: p.grad.float() for k, p in net.named_parameters()}, dx=inp.grad.float(), out=out.float())
    close(res["tc"]["out"], fx[f"{name}.out"], rtol=2e-2, atol=2e-2, msg=f"{name} out vs reference")
    for k, v in res["tc"].items():
        r = res["direct"][k]
        close(v, r, rtol=1e-2, atol=1e-2 * max(1e-9, r.abs().max().item()), msg=f"{name} {k}: tcgen05 vs direct")
