"""GPU parity of the fused BatchNorm2d + activation + Dropout2d op (csrc/norm_act_kernels.cu, cgat.norm_act) -- the
non-conv part of the reference's ConvBlock (dcgan/model.py:35-52) and of the SmaAt-UNet double convs -- against
torch.nn.BatchNorm2d / the torch activations on the CPU (fp32), forward, backward and running statistics."""
import pytest
import torch
import torch.nn.functional as F

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_act(z, act):
    return {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.2), 3: torch.sigmoid}[act](z)


@pytest.mark.parametrize("C,shape", [(64, (5, 9, 7)), (8, (3, 16, 16)), (6, (4, 5, 5)), (1, (2, 8, 8)), (512, (2, 4, 4))])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_act_fp32_matches_torch(C, shape, act, training):
    from cgat.norm_act import BatchNormAct2d

    N, H, W = shape
    torch.manual_seed(C + act)
    ref = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
        ref.running_mean.uniform_(-0.2, 0.2)
        ref.running_var.uniform_(0.5, 1.5)
    ours = BatchNormAct2d(C, act=act, slope=0.2)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV)
    ref.train(training)
    ours.train(training)
    x = torch.randn(N, C, H, W) * 1.5 + 0.3
    g = torch.randn(N, C, H, W)
    xr = x.clone().requires_grad_()
    yr = _ref_act(ref(xr), act)
    yr.backward(g)
    xo = x.to(DEV).requires_grad_()
    yo = ours(xo)
    yo.backward(g.to(DEV))
    torch.cuda.synchronize()
    close(yo, yr.detach(), rtol=1e-4, atol=1e-5, msg="y")
    close(xo.grad, xr.grad, rtol=1e-4, atol=2e-5, msg="dx")
    close(ours.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-4, msg="dgamma")
    close(ours.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4, msg="dbeta")
    for k, v in ours.state_dict().items():  # running statistics and num_batches_tracked move exactly as torch's
        close(v, ref.state_dict()[k], rtol=1e-5, atol=1e-6, msg=k)


def test_batchnorm_act_bf16_and_state_dict_keys():
    from cgat.norm_act import ACT_LRELU, BatchNormAct2d

    torch.manual_seed(2)
    ours = BatchNormAct2d(32, act=ACT_LRELU).to(DEV).train()
    assert list(ours.state_dict().keys()) == list(torch.nn.BatchNorm2d(32).state_dict().keys())
    x = torch.randn(6, 32, 10, 12).bfloat16()
    ref = F.leaky_relu(F.batch_norm(x.float(), None, None, training=True), 0.2)
    y = ours(x.to(DEV).contiguous(memory_format=torch.channels_last))
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    close(y, ref, rtol=2e-2, atol=2e-2, msg="bf16 y")


def test_dropout2d_mask_is_channelwise_scaled_and_replayable():
    """Dropout2d (dcgan/model.py:46-47, p = 0.01 in every ConvBlock): whole channels of a sample are zeroed, survivors
    scaled by 1/(1-p); a fresh mask per call, the same sequence for the same seed."""
    from cgat.norm_act import ACT_RELU, ActDropout2d

    def run(seed, p, calls):
        torch.manual_seed(seed)
        m = ActDropout2d(act=ACT_RELU, dropout=p).to(DEV).train()
        x = torch.ones(64, 128, 4, 4, device=DEV)
        return [m(x) for _ in range(calls)]

    p = 0.25
    ys = run(5, p, 3)
    for y in ys:
        per = y.amax(dim=(2, 3))
        assert torch.equal(per, y.amin(dim=(2, 3))), "a channel of a sample is dropped or kept as a whole"
        vals = set(per.unique().tolist())
        assert all(v == 0.0 or abs(v - 1 / (1 - p)) < 1e-5 for v in vals), vals
        keep = float((per > 0).float().mean())
        assert abs(keep - (1 - p)) < 0.02, keep  # 8192 Bernoulli draws: sigma = 0.005
    assert not torch.equal(ys[0], ys[1]) and not torch.equal(ys[1], ys[2]), "every call draws a new mask"
    again = run(5, p, 3)
    assert all(torch.equal(a, b) for a, b in zip(ys, again)), "same seed, same sequence"
    m = ActDropout2d(act=ACT_RELU, dropout=p).to(DEV).eval()
    x = torch.randn(2, 8, 3, 3, device=DEV)
    assert torch.equal(m(x), F.relu(x)), "eval mode: no dropout"


def test_dropout_before_sigmoid_keeps_the_block_order():
    """ConvBlock(batchnorm=False, act=sigmoid) (the generator's last block): a dropped channel is sigmoid(0) = 0.5."""
    from cgat.norm_act import ACT_SIGMOID, ActDropout2d

    torch.manual_seed(1)
    m = ActDropout2d(act=ACT_SIGMOID, dropout=0.5).to(DEV).train()
    x = torch.randn(16, 32, 4, 4, device=DEV).requires_grad_()
    y = m(x)
    y.sum().backward()
    dropped = (y == 0.5).all(dim=3).all(dim=2)
    assert 0.3 < float(dropped.float().mean()) < 0.7
    want = torch.sigmoid(2.0 * x.detach())  # kept channels: sigmoid(x / (1 - p))
    assert torch.allclose(y.detach()[~dropped], want[~dropped], rtol=1e-5, atol=1e-6)
    assert (x.grad[dropped] == 0).all()


@pytest.mark.parametrize("C,shape,sets", [(64, (6, 9, 7), 3), (8, (8, 16, 16), 8), (1, (4, 8, 8), 2), (512, (4, 4, 4), 2)])
@pytest.mark.parametrize("act", [1, 3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_batchnorm_statistic_sets_match_sequential_torch_calls(C, shape, sets, act, dtype):
    """``sets`` groups of the batch through ONE launch per pass == the reference's UnetModel calling the same BatchNorm2d
    once per vertex (unet_model.py:25-26): per-group statistics, `sets` running-stat updates in group order, parameter
    gradients summed over the groups."""
    from cgat.norm_act import BatchNormAct2d

    N, H, W = shape
    torch.manual_seed(C + act + sets)
    ref = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
        ref.running_mean.uniform_(-0.2, 0.2)
        ref.running_var.uniform_(0.5, 1.5)
    ours = BatchNormAct2d(C, act=act, slope=0.2)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).train()
    ours.sets = sets
    ref.train()
    x = (torch.randn(N, C, H, W) * 1.5 + 0.3).to(dtype).float()
    x = x + torch.arange(sets).repeat_interleave(N // sets).view(N, 1, 1, 1) * 0.7  # the groups differ in mean
    x = x.to(dtype).float()
    g = torch.randn(N, C, H, W).to(dtype).float()
    xr = x.clone().requires_grad_()
    nb = N // sets
    yr = torch.cat([_ref_act(ref(xr[i * nb:(i + 1) * nb]), act) for i in range(sets)])
    yr.backward(g)
    xo = x.to(DEV).to(dtype).requires_grad_()
    yo = ours(xo)
    yo.backward(g.to(DEV).to(dtype))
    torch.cuda.synchronize()
    lo = dtype == torch.bfloat16
    close(yo.float(), yr.detach(), rtol=2e-2 if lo else 1e-4, atol=2e-2 if lo else 1e-5, msg="y")
    close(xo.grad.float(), xr.grad, rtol=2e-2 if lo else 1e-4, atol=3e-2 if lo else 2e-5, msg="dx")
    close(ours.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-4 * N * H * W if lo else 1e-4, msg="dgamma")
    close(ours.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4 * N * H * W if lo else 1e-4, msg="dbeta")
    for k, v in ours.state_dict().items():
        close(v, ref.state_dict()[k], rtol=1e-5, atol=1e-6, msg=k)
    assert int(ours.num_batches_tracked) == sets
