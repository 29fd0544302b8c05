"""GPU parity of the small-channel wgrad kernel (csrc/conv_wgrad_small.cu: the DCGAN generator's k = 4 ``padding="same"``
convs on 64 x 64 frames, dcgan/model.py:55-76) vs torch CPU conv2d: dw and dbias, fp32 (rtol 1e-4) and bf16."""
import pytest
import torch
import torch.nn.functional as F

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # n, h, w, cin, cout, k, pad(t,l,b,r)
    (2, 64, 64, 32, 16, 4, (1, 1, 2, 2)),   # 512 (tap, ci) pairs: two per thread
    (2, 64, 64, 16, 8, 4, (1, 1, 2, 2)),    # 256 pairs: one per thread
    (2, 64, 64, 8, 4, 4, (1, 1, 2, 2)),     # 128 pairs: two pixel groups
    (2, 64, 64, 4, 32, 4, (1, 1, 2, 2)),    # first generator layer: 64 pairs, 32 output channels in registers
    (2, 64, 64, 4, 4, 4, (1, 1, 2, 2)),     # last generator layer
    (3, 45, 37, 8, 8, 3, (1, 1, 1, 1)),     # ragged tiles, 72 pairs (256 % 72 != 0: idle tail threads)
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
def test_wgrad_small(case, dtype):
    from cgat.functional import IMPL_DIRECT, conv2d_nhwc

    n, h, w, cin, cout, k, pad = case
    torch.manual_seed(17)
    x = (torch.rand(n, h, w, cin) - 0.5).to(dtype).float()
    wt = (torch.rand(cout, k, k, cin) - 0.5).to(dtype).float()
    b = torch.rand(cout) - 0.5
    xr, wr, br = (t.clone().requires_grad_() for t in (x, wt, b))
    pt, pl, pb, pr = pad
    yr = F.conv2d(F.pad(xr.permute(0, 3, 1, 2), (pl, pr, pt, pb)), wr.permute(0, 3, 1, 2), br).permute(0, 2, 3, 1)
    g = (torch.rand_like(yr) - 0.5).to(dtype).float()
    yr.backward(g)
    xo = x.to(DEV, dtype).requires_grad_()
    wo = wt.to(DEV).requires_grad_()
    bo = b.to(DEV).requires_grad_()
    yo = conv2d_nhwc(xo, wo, bo, stride=1, pad=pad, impl=IMPL_DIRECT)
    yo.backward(g.to(DEV, dtype))
    tol = 1e-4 if dtype == torch.float32 else 1e-3  # same bf16-rounded operands, fp32 accumulation
    close(wo.grad, wr.grad, rtol=tol, atol=tol * max(1.0, wr.grad.abs().max().item()), msg="dw")
    close(bo.grad, br.grad, rtol=tol, atol=tol * max(1.0, br.grad.abs().max().item()), msg="db")
