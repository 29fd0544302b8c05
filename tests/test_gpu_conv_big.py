"""GPU parity of the streamed-operand tcgen05 conv kernels (csrc/conv_tc_big.cu: large channel counts, bf16) vs torch
CPU conv2d on bf16-rounded inputs: fprop, dgrad, wgrad (split-K) and dbias through the C ABI."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

BIG_CASES = [
    # n, h, w, cin, cout, k, pad(t,l,b,r)
    (2, 16, 16, 64, 64, 1, (0, 0, 0, 0)),       # pointwise, one 64-channel block
    (2, 17, 17, 256, 128, 2, (0, 0, 0, 0)),     # DCGAN discriminator conv2 after 2x2 regrouping (dcgan/model.py:152)
    (4, 9, 9, 512, 256, 2, (0, 0, 0, 0)),       # conv3: 8x8 outputs, two images per pixel tile
    (16, 5, 5, 1024, 512, 2, (0, 0, 0, 0)),     # conv4: 4x4 outputs, eight images per tile, two cout tiles
    (2, 12, 10, 128, 72, 3, (1, 1, 1, 1)),      # ragged pixel tiles, cout not a multiple of 16
    (3, 8, 8, 72, 200, 3, (1, 1, 1, 1)),        # channel blocks that run past cin / cout (TMA zero fill)
    (2, 16, 16, 64, 384, 1, (0, 0, 0, 0)),      # second cout tile half empty
    (2, 16, 16, 64, 64, 4, (1, 1, 2, 2)),       # k=4 padding="same" (dcgan/model.py:61-72)
    (1, 3, 3, 64, 64, 3, (1, 1, 1, 1)),         # image smaller than any tile
    (1, 20, 36, 192, 640, 3, (1, 1, 1, 1)),     # three cout tiles (the last half empty), 3 channel blocks, ragged rows
    (130, 5, 5, 128, 64, 2, (0, 0, 0, 0)),      # 4x4 outputs, 8 images per pixel tile, batch not a multiple of 8
    (2, 16, 16, 32, 16, 4, (1, 1, 2, 2)),       # DCGAN generator layers (dcgan/model.py:61-72) on few pixels: fprop/dgrad
    (2, 16, 16, 16, 8, 4, (1, 1, 2, 2)),        #   on the resident-weight kernels, wgrad on the streamed one (one dY box)
]


def _ref_conv(x_nhwc, w_krsc, bias, pad):
    x = x_nhwc.permute(0, 3, 1, 2)
    w = w_krsc.permute(0, 3, 1, 2)
    pt, pl, pb, pr = pad
    y = F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, bias, 1)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("case", BIG_CASES)
def test_conv_big_fprop_dgrad_wgrad(case):
    from cgat import _lib
    from cgat.functional import IMPL_TC, _conv_desc, conv2d_nhwc

    n, h, w, cin, cout, k, pad = case
    torch.manual_seed(11)
    x = (torch.rand(n, h, w, cin) - 0.5).bfloat16().float()
    wt = (torch.rand(cout, k, k, cin) - 0.5).bfloat16().float()
    b = torch.rand(cout) - 0.5
    xr, wr, br = (t.clone().requires_grad_() for t in (x, wt, b))
    yr = _ref_conv(xr, wr, br, pad)
    g = (torch.rand_like(yr) - 0.5).bfloat16().float()
    yr.backward(g)
    d = _conv_desc(n, h, w, cin, cout, k, k, 1, pad[0], pad[1], yr.shape[1], yr.shape[2], _lib.BF16, 0)
    for which in range(3):
        assert _lib.lib().cgat_conv_tc_supported(ctypes.byref(d), which) == 1
    xo = x.to(DEV, torch.bfloat16).requires_grad_()
    wo = wt.to(DEV).requires_grad_()
    bo = b.to(DEV).requires_grad_()
    yo = conv2d_nhwc(xo, wo, bo, stride=1, pad=pad, impl=IMPL_TC)
    scale = max(1.0, yr.abs().max().item())
    close(yo, yr.detach(), rtol=2e-2, atol=1e-2 * scale, msg="y")
    yo.backward(g.to(DEV, torch.bfloat16))
    close(xo.grad, xr.grad, rtol=2e-2, atol=1e-2 * max(1.0, xr.grad.abs().max().item()), msg="dx")
    # wgrad accumulates in fp32 from the same bf16 operands: tight
    close(wo.grad, wr.grad, rtol=1e-3, atol=1e-3 * max(1.0, wr.grad.abs().max().item()), msg="dw")
    close(bo.grad, br.grad, rtol=1e-3, atol=1e-3 * max(1.0, br.grad.abs().max().item()), msg="db")


def test_conv_big_activation_epilogue():
    from cgat import _lib
    from cgat.functional import IMPL_TC, conv2d_nhwc

    torch.manual_seed(3)
    x = (torch.rand(2, 8, 8, 128) - 0.5).bfloat16()
    wt = ((torch.rand(64, 1, 1, 128) - 0.5) * 0.2).bfloat16()
    b = torch.rand(64) - 0.5
    z = _ref_conv(x.float(), wt.float(), b, (0, 0, 0, 0))
    for act, fn in ((_lib.ACT_RELU, torch.relu), (_lib.ACT_LRELU, lambda t: F.leaky_relu(t, 0.2)),
                    (_lib.ACT_SIGMOID, torch.sigmoid)):
        y = conv2d_nhwc(x.to(DEV), wt.to(DEV), b.to(DEV), stride=1, pad=(0, 0, 0, 0), act=act, impl=IMPL_TC)
        close(y, fn(z), rtol=2e-2, atol=1e-2, msg=f"act {act}")


def test_dcgan_discriminator_conv_routes_to_tensor_cores():
    """The k=4 s=2 p=1 convs of FrameDiscriminator (dcgan/model.py:152-160) at the reference's ndf=64 take the 2x2
    regrouping + streamed tcgen05 kernels in all three directions, and match cuDNN-free fp32 math."""
    from cgat.conv_layers import Conv2d

    torch.manual_seed(5)
    conv = Conv2d(128, 256, 4, 2, 1, bias=False).to(DEV)
    x = (torch.rand(4, 128, 16, 16) - 0.5).bfloat16()
    xr = x.float().requires_grad_()
    wr = conv.weight.detach().cpu().bfloat16().float().requires_grad_()
    yr = F.conv2d(xr, wr, None, 2, 1)
    g = (torch.rand_like(yr) - 0.5).bfloat16().float()
    yr.backward(g)
    xo = x.to(DEV).requires_grad_()
    assert conv._space_to_depth_route(xo.permute(0, 2, 3, 1)) is not None
    with torch.autocast("cuda", enabled=False):
        yo = conv(xo)
    yo.backward(g.to(DEV, torch.bfloat16))
    close(yo, yr.detach(), rtol=2e-2, atol=1e-2 * max(1.0, yr.abs().max().item()), msg="y")
    close(xo.grad, xr.grad, rtol=2e-2, atol=1e-2 * max(1.0, xr.grad.abs().max().item()), msg="dx")
    close(conv.weight.grad, wr.grad, rtol=2e-2, atol=1e-2 * max(1.0, wr.grad.abs().max().item()), msg="dw")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("stride", [1, 4])
def test_fullwindow_conv(dtype, stride):
    """Last conv of the DCGAN discriminators (dcgan/model.py:166-169, 512 -> 1, k=4 over the 4x4 map; the temporal one
    with stride 4): one dot product per image through the full-window kernel."""
    from cgat.conv_layers import Conv2d

    torch.manual_seed(9)
    conv = Conv2d(512, 1, 4, stride, 0, bias=False).to(DEV)
    x = (torch.rand(6, 512, 4, 4) - 0.5).to(dtype)
    wr = conv.weight.detach().cpu().to(dtype).float().requires_grad_()
    xr = x.float().clone().requires_grad_()
    yr = F.conv2d(xr, wr, None, stride, 0)
    yr.backward(torch.ones_like(yr))
    xo = x.clone().to(DEV).requires_grad_()
    yo = conv(xo)
    assert yo.shape == yr.shape
    yo.backward(torch.ones_like(yo))
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    close(yo, yr.detach(), rtol=tol, atol=tol, msg="y")
    close(xo.grad, xr.grad, rtol=tol, atol=tol, msg="dx")
    close(conv.weight.grad, wr.grad, rtol=tol, atol=tol, msg="dw")


def test_conv_big_rectangular_kernel():
    """1x3 and 3x1 kernels (kh != kw) through the tap-shifted TMA boxes: fprop, dgrad, wgrad."""
    from cgat.functional import IMPL_TC, conv2d_nhwc

    torch.manual_seed(21)
    for kh, kw, pad in ((1, 3, (0, 1, 0, 1)), (3, 1, (1, 0, 1, 0))):
        x = (torch.rand(2, 12, 14, 64) - 0.5).bfloat16().float()
        wt = (torch.rand(96, kh, kw, 64) - 0.5).bfloat16().float()
        xr, wr = x.clone().requires_grad_(), wt.clone().requires_grad_()
        pt, pl, pb, pr = pad
        yr = F.conv2d(F.pad(xr.permute(0, 3, 1, 2), (pl, pr, pt, pb)), wr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        g = (torch.rand_like(yr) - 0.5).bfloat16().float()
        yr.backward(g)
        xo = x.to(DEV, torch.bfloat16).requires_grad_()
        wo = wt.to(DEV).requires_grad_()
        yo = conv2d_nhwc(xo, wo, None, stride=1, pad=pad, impl=IMPL_TC)
        yo.backward(g.to(DEV, torch.bfloat16))
        close(yo, yr.detach(), rtol=2e-2, atol=1e-2 * max(1.0, yr.abs().max().item()), msg=f"y {kh}x{kw}")
        close(xo.grad, xr.grad, rtol=2e-2, atol=1e-2 * max(1.0, xr.grad.abs().max().item()), msg=f"dx {kh}x{kw}")
        close(wo.grad, wr.grad, rtol=1e-3, atol=1e-3 * max(1.0, wr.grad.abs().max().item()), msg=f"dw {kh}x{kw}")
