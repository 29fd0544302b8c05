"""GPU parity: direct NHWC conv kernels vs torch CPU conv2d; fused loss and Adam vs the oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import spec
from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad(t,l,b,r)
    (2, 12, 10, 24, 72, 3, 1, (1, 1, 1, 1)),   # conv-GAT dense node conv
    (2, 16, 16, 4, 32, 4, 1, (1, 1, 2, 2)),    # DCGAN generator: k=4 padding="same" (asymmetric)
    (2, 16, 16, 8, 16, 4, 2, (1, 1, 1, 1)),    # discriminator stride-2
    (2, 4, 4, 16, 1, 4, 1, (0, 0, 0, 0)),      # FrameDiscriminator conv5
    (2, 8, 8, 16, 1, 4, 4, (0, 0, 0, 0)),      # TemporalDiscriminator last conv, stride 4
    (1, 7, 5, 3, 5, 1, 1, (0, 0, 0, 0)),       # pointwise, odd sizes
]


def _ref_conv(x_nhwc, w_krsc, bias, stride, pad):
    x = x_nhwc.permute(0, 3, 1, 2)
    w = w_krsc.permute(0, 3, 1, 2)
    pt, pl, pb, pr = pad
    y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, bias, stride)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_direct_fprop_dgrad_wgrad(case, dtype):
    from cgat.functional import IMPL_DIRECT, conv2d_nhwc

    n, h, w, cin, cout, k, stride, pad = case
    torch.manual_seed(5)
    x = (torch.rand(n, h, w, cin) - 0.5).to(dtype).float()
    wt = (torch.rand(cout, k, k, cin) - 0.5).to(dtype).float()
    b = torch.rand(cout) - 0.5
    xr, wr, br = (t.clone().requires_grad_() for t in (x, wt, b))
    yr = _ref_conv(xr, wr, br, stride, pad)
    g = (torch.rand_like(yr) - 0.5).to(dtype).float()
    yr.backward(g)
    xo = x.to(DEV, dtype).requires_grad_()
    wo = wt.to(DEV).requires_grad_()
    bo = b.to(DEV).requires_grad_()
    yo = conv2d_nhwc(xo, wo, bo, stride=stride, pad=pad, impl=IMPL_DIRECT)
    rtol, atol = (1e-4, 1e-5) if dtype == torch.float32 else (2e-2, 2e-2)
    close(yo, yr.detach(), rtol=rtol, atol=atol, msg="y")
    yo.backward(g.to(DEV, dtype))
    close(xo.grad, xr.grad, rtol=rtol, atol=atol, msg="dx")
    close(wo.grad, wr.grad, rtol=rtol, atol=atol * max(1.0, wr.grad.abs().max().item()), msg="dw")
    close(bo.grad, br.grad, rtol=rtol, atol=atol * max(1.0, br.grad.abs().max().item()), msg="db")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_loss_and_gradient(dtype):
    from cgat.functional import loss_and_grad

    torch.manual_seed(1)
    yh = torch.rand(3, 9, 7, 4, 6).to(dtype).float()
    y = torch.rand(3, 9, 7, 4, 6).to(dtype).float()
    yr = yh.clone().requires_grad_()
    lr = spec.train_loss(yr, y)  # convolutional_gat/train.py:131
    lr.backward()
    loss, dy = loss_and_grad(yh.to(DEV, dtype), y.to(DEV, dtype))
    tol = dict(rtol=1e-5, atol=1e-7) if dtype == torch.float32 else dict(rtol=2e-2, atol=1e-6)
    close(loss[0], lr.detach(), rtol=1e-5, atol=1e-6, msg="loss")
    close(dy, yr.grad, msg="dloss", **tol)


def test_fused_adam_matches_torch_optim():
    from cgat.functional import adam_step_

    torch.manual_seed(2)
    p0 = torch.rand(1000) - 0.5
    pr = p0.clone().requires_grad_()
    opt = torch.optim.Adam([pr], lr=1e-3, weight_decay=0.01)  # convolutional_gat/train.py:212
    p = p0.to(DEV)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    for it in range(5):
        g = torch.rand(1000) - 0.5
        pr.grad = g.clone()
        opt.step()
        step += 1
        adam_step_(p, g.to(DEV), m, v, step, 1e-3)
        close(p, pr.detach(), rtol=1e-5, atol=1e-7, msg=f"param after step {it + 1}")
        po, mo, vo = spec.adam_step(p0 if it == 0 else po, g, torch.zeros(1000) if it == 0 else mo,
                                    torch.zeros(1000) if it == 0 else vo, it + 1, 1e-3)
        close(p, po, rtol=1e-5, atol=1e-7, msg="oracle adam")
