"""GPU parity of the SmaAt-UNet's non-conv ops (csrc/unet_glue_kernels.cu, cgat.unet_ops) -- max-pooling, bilinear
up-sampling + pad + concat, CBAM's channel and spatial gates -- against the torch ops the public architecture is written
in (convolutional_gat/unet_model.py:20 builds that net), forward and backward, fp32 at 1e-4 and bf16 at 2e-2."""
import pytest
import torch
import torch.nn.functional as F

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(fn_ref, fn_ours, inputs, params=(), rtol=1e-4, atol=1e-5, dtype=torch.float32):
    ins_r = [t.clone().requires_grad_() for t in inputs]
    par_r = [p.clone().requires_grad_() for p in params]
    out_r = fn_ref(*ins_r, *par_r)
    g = torch.randn_like(out_r)
    out_r.backward(g)
    ins_o = [t.to(DEV).to(dtype).requires_grad_() for t in inputs]
    par_o = [p.to(DEV).requires_grad_() for p in params]
    out_o = fn_ours(*ins_o, *par_o)
    out_o.backward(g.to(DEV).to(out_o.dtype))
    torch.cuda.synchronize()
    close(out_o, out_r.detach(), rtol=rtol, atol=atol, msg="out")
    for i, (a, b) in enumerate(zip(ins_o, ins_r)):
        close(a.grad, b.grad, rtol=rtol, atol=atol * 4, msg=f"d(input {i})")
    for i, (a, b) in enumerate(zip(par_o, par_r)):
        close(a.grad, b.grad, rtol=rtol * 3, atol=atol * 20, msg=f"d(param {i})")


@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (3, 8, 9, 7), (2, 6, 5, 4), (1, 512, 4, 4)])
def test_maxpool2(shape):
    from cgat.unet_ops import MaxPool2d

    torch.manual_seed(1)
    _run(lambda x: F.max_pool2d(x, 2), MaxPool2d(), [torch.randn(*shape)])


@pytest.mark.parametrize("n,c1,h1,w1,c2,h,w", [(2, 64, 8, 8, 64, 16, 16), (2, 8, 5, 6, 16, 11, 13), (1, 6, 3, 3, 4, 7, 6),
                                               (2, 512, 1, 1, 512, 2, 2)])
def test_upsample_pad_concat(n, c1, h1, w1, c2, h, w):
    from cgat.unet_ops import upsample_pad_concat

    def ref(x1, x2):  # UpDS.forward of the public SmaAt-UNet
        x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
        dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
        x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return torch.cat([x2, x1], dim=1)

    torch.manual_seed(2)
    _run(ref, upsample_pad_concat, [torch.randn(n, c1, h1, w1), torch.randn(n, c2, h, w)])


@pytest.mark.parametrize("n,c,h,w", [(2, 64, 16, 16), (3, 32, 7, 9), (2, 512, 4, 4), (1, 128, 40, 40)])
def test_cbam_channel_gate(n, c, h, w):
    from cgat.unet_ops import channel_gate

    hid = c // 16

    def ref(x, w1, b1, w2, b2):
        mlp = lambda v: F.linear(F.relu(F.linear(v.flatten(1), w1, b1)), w2, b2)
        s = mlp(F.adaptive_avg_pool2d(x, 1)) + mlp(F.adaptive_max_pool2d(x, 1))
        return x * torch.sigmoid(s)[:, :, None, None]

    torch.manual_seed(3)
    params = [torch.randn(hid, c) * 0.2, torch.randn(hid) * 0.2, torch.randn(c, hid) * 0.2, torch.randn(c) * 0.2]
    _run(ref, channel_gate, [torch.randn(n, c, h, w)], params)


@pytest.mark.parametrize("n,c,h,w", [(2, 64, 16, 16), (3, 12, 7, 9), (2, 512, 4, 4)])
def test_cbam_spatial_pool_and_gate(n, c, h, w):
    from cgat.unet_ops import channel_pool, pixel_gate

    torch.manual_seed(4)
    _run(lambda x: torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1), channel_pool, [torch.randn(n, c, h, w)])
    _run(lambda x, s: x * s, pixel_gate, [torch.randn(n, c, h, w), torch.rand(n, 1, h, w)])


def test_unet_ops_bf16():
    from cgat.unet_ops import MaxPool2d, channel_gate, upsample_pad_concat

    torch.manual_seed(5)
    x = torch.randn(2, 64, 16, 16).bfloat16().float()
    _run(lambda t: F.max_pool2d(t, 2), MaxPool2d(), [x], rtol=2e-2, atol=2e-2, dtype=torch.bfloat16)
    x1, x2 = torch.randn(2, 32, 8, 8).bfloat16().float(), torch.randn(2, 16, 17, 16).bfloat16().float()

    def ref(a, b):
        a = F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=True)
        dy, dx = b.shape[2] - a.shape[2], b.shape[3] - a.shape[3]
        return torch.cat([b, F.pad(a, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])], dim=1)

    _run(ref, upsample_pad_concat, [x1, x2], rtol=2e-2, atol=2e-2, dtype=torch.bfloat16)
    c, hid = 64, 4
    params = [torch.randn(hid, c) * 0.2, torch.randn(hid) * 0.2, torch.randn(c, hid) * 0.2, torch.randn(c) * 0.2]

    def refg(t, w1, b1, w2, b2):
        mlp = lambda v: F.linear(F.relu(F.linear(v.flatten(1), w1, b1)), w2, b2)
        return t * torch.sigmoid(mlp(F.adaptive_avg_pool2d(t, 1)) + mlp(F.adaptive_max_pool2d(t, 1)))[:, :, None, None]

    _run(refg, channel_gate, [x], params, rtol=2e-2, atol=3e-2, dtype=torch.bfloat16)


def test_smaat_unet_train_and_eval_passes_use_no_torch_glue():
    """Whole SmaAt-UNet, forward + backward: every op between input and loss is a C-ABI call of ours (cgat._lib's profile
    lists them).  Train mode: output and BatchNorm running statistics against the oracle module (oracle/spec.py).  The
    parameter gradients are compared in eval mode: in train mode the bottleneck BatchNorm of a 32x32 input normalises
    over 16 values and amplifies fp32 rounding differences to per cents (every op's own backward is pinned above)."""
    from cgat import _lib
    from convolutional_gat.GAT3D.smaat_unet.SmaAt_UNet import SmaAt_UNet
    from oracle import spec

    torch.manual_seed(7)
    ref = spec.SpecSmaAtUNet(4, 4).train()
    ours = SmaAt_UNet(4, 4)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).train()
    x = torch.rand(4, 4, 32, 32)
    out_r = ref(x)
    g = torch.randn_like(out_r)
    _lib.profile_start()
    out_o = ours(x.to(DEV))
    out_o.backward(g.to(DEV))
    torch.cuda.synchronize()
    prof = _lib.profile_stop()
    for name in ("cgat_maxpool2_fwd", "cgat_maxpool2_bwd", "cgat_upcat_fwd", "cgat_upcat_bwd", "cgat_pool_hw", "cgat_dot_hw",
                 "cgat_cbam_mlp_fwd", "cgat_cbam_mlp_bwd", "cgat_gate_channels_fwd", "cgat_gate_channels_bwd",
                 "cgat_chan_pool_fwd", "cgat_chan_pool_bwd", "cgat_gate_pixels", "cgat_chan_dot", "cgat_bn_stats_sets",
                 "cgat_bn_act_fwd_sets", "cgat_bn_act_bwd_sets"):
        assert name in prof, (name, sorted(prof))
    close(out_o, out_r.detach(), rtol=1e-3, atol=1e-4, msg="SmaAt-UNet train-mode out")
    sr = ref.state_dict()
    for k, v in ours.state_dict().items():
        if "running" in k or "num_batches" in k:
            close(v, sr[k], rtol=1e-4, atol=1e-5, msg=k)
    ref.eval()
    ours.eval()
    for p in list(ref.parameters()) + list(ours.parameters()):
        p.grad = None
    out_r = ref(x)
    out_r.backward(g)
    out_o = ours(x.to(DEV))
    out_o.backward(g.to(DEV))
    torch.cuda.synchronize()
    close(out_o, out_r.detach(), rtol=1e-3, atol=1e-4, msg="SmaAt-UNet eval-mode out")
    pr = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        gr = pr[k].grad
        close(p.grad, gr, rtol=2e-3, atol=2e-4 * max(1e-6, float(gr.abs().max())), msg=f"d{k}")
