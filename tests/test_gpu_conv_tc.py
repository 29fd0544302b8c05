"""GPU parity of the tcgen05 implicit-GEMM conv kernels (bf16) vs torch CPU conv2d on bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

TC_CASES = [
    # n, h, w, cin, cout, k, pad(t,l,b,r)
    (1, 16, 8, 16, 16, 1, (0, 0, 0, 0)),     # one tile, pointwise: plain GEMM through the descriptors
    (2, 16, 16, 24, 16, 1, (0, 0, 0, 0)),    # odd chunk count -> zero padding chunk
    (2, 16, 16, 16, 32, 3, (1, 1, 1, 1)),    # 3x3: tap-shifted descriptors into the halo tile
    (2, 13, 11, 24, 72, 3, (1, 1, 1, 1)),    # conv-GAT dense node conv, ragged tiles
    (3, 20, 20, 24, 72, 3, (1, 1, 1, 1)),    # KNMI 20x20 crop
    (2, 16, 16, 32, 16, 4, (1, 1, 2, 2)),    # DCGAN generator layer: k=4 padding="same"
    (2, 9, 9, 64, 128, 1, (0, 0, 0, 0)),     # UNet-like pointwise
    (1, 40, 24, 8, 8, 3, (1, 1, 1, 1)),      # minimum channels
]


def _ref_conv(x_nhwc, w_krsc, bias, pad):
    x = x_nhwc.permute(0, 3, 1, 2)
    w = w_krsc.permute(0, 3, 1, 2)
    pt, pl, pb, pr = pad
    y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, bias, 1)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_fprop_dgrad(case):
    from cgat.functional import IMPL_AUTO, IMPL_TC, conv2d_nhwc

    n, h, w, cin, cout, k, pad = case
    torch.manual_seed(7)
    x = (torch.rand(n, h, w, cin) - 0.5).bfloat16().float()
    wt = (torch.rand(cout, k, k, cin) - 0.5).bfloat16().float()
    b = torch.rand(cout) - 0.5
    xr, wr, br = (t.clone().requires_grad_() for t in (x, wt, b))
    yr = _ref_conv(xr, wr, br, pad)
    g = (torch.rand_like(yr) - 0.5).bfloat16().float()
    yr.backward(g)
    xo = x.to(DEV, torch.bfloat16).requires_grad_()
    wo = wt.to(DEV).requires_grad_()
    bo = b.to(DEV).requires_grad_()
    wgrad_cols = k * k * ((cin // 8 + 1) // 2 * 16) + 16  # TMEM columns the wgrad accumulators need
    tc_bwd = cout % 8 == 0 and cout <= 128 and wgrad_cols <= 512
    # shapes the resident-weight wgrad does not take fall to the streamed one (conv_tc_big.cu) from 16 x 8 channels up
    # (unless the CUDA-core small-channel wgrad takes them: <= 32 channels and at least 4096 output pixels)
    small = cin <= 32 and cout in (4, 8, 16, 32) and k * k * cin <= 512 and n * yr.shape[1] * yr.shape[2] >= 4096
    tc_bwd = tc_bwd or (cin % 8 == 0 and cout % 8 == 0 and cin >= 16 and cout >= 8 and not small)
    # IMPL_AUTO falls back to the direct wgrad kernel where the tcgen05 one does not serve the shape
    yo = conv2d_nhwc(xo, wo, bo, stride=1, pad=pad, impl=IMPL_TC if tc_bwd else IMPL_AUTO)
    scale = max(1.0, yr.abs().max().item())
    close(yo, yr.detach(), rtol=2e-2, atol=1e-2 * scale, msg="y")
    # full backward through tcgen05: dgrad (K2) and wgrad + dbias (K3)
    import ctypes

    from cgat import _lib
    from cgat.functional import _conv_desc

    ho, wo_ = yr.shape[1], yr.shape[2]
    d = _conv_desc(n, h, w, cin, cout, k, k, 1, pad[0], pad[1], ho, wo_, _lib.BF16, 0)
    assert _lib.lib().cgat_conv_tc_supported(ctypes.byref(d), 1) == 1
    assert _lib.lib().cgat_conv_tc_supported(ctypes.byref(d), 0) == 1
    assert _lib.lib().cgat_conv_tc_supported(ctypes.byref(d), 2) == int(tc_bwd)
    yo.backward(g.to(DEV, torch.bfloat16))
    gscale = max(1.0, xr.grad.abs().max().item())
    close(xo.grad, xr.grad, rtol=2e-2, atol=1e-2 * gscale, msg="dx")
    close(wo.grad, wr.grad, rtol=2e-2, atol=1e-2 * max(1.0, wr.grad.abs().max().item()), msg="dw")
    close(bo.grad, br.grad, rtol=2e-2, atol=1e-2 * max(1.0, br.grad.abs().max().item()), msg="db")
