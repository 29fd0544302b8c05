"""GPU parity of the small-channel direct conv kernel (csrc/conv_small.cu: the ends of the DCGAN generator, k = 4
``padding="same"`` convs between 4 / 8 and up to 32 channels on 64 x 64 frames, dcgan/model.py:19-34) vs torch CPU conv2d:
fprop with bias + activation, and dgrad (the same kernel over dy, taps flipped, padding mirrored)."""
import pytest
import torch
import torch.nn.functional as F

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # n, h, w, cin, cout, k, pad(t,l,b,r), act
    (2, 64, 64, 4, 32, 4, (1, 1, 2, 2), 1),   # first generator layer (ReLU)
    (2, 64, 64, 4, 4, 4, (1, 1, 2, 2), 3),    # last generator layer (sigmoid)
    (2, 64, 64, 8, 4, 4, (1, 1, 2, 2), 0),    # dgrad of 8 -> 4 contracts over 4 channels
    (3, 45, 37, 4, 8, 3, (1, 1, 1, 1), 2),    # ragged rows (37 % 4 != 0), LeakyReLU
    (2, 19, 23, 8, 16, 5, (2, 2, 2, 2), 0),
    (1, 9, 7, 4, 16, 1, (0, 0, 0, 0), 0),
    (3, 32, 32, 2, 1, 7, (3, 3, 3, 3), 0),    # CBAM spatial gate conv (no bias in the model; the kernel takes one)
    (2, 17, 13, 2, 1, 7, (3, 3, 3, 3), 0),
]


def _act(t, act):
    return {0: lambda v: v, 1: F.relu, 2: lambda v: F.leaky_relu(v, 0.2), 3: torch.sigmoid}[act](t)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
def test_conv_small_fprop_and_dgrad(case, dtype):
    from cgat import _lib
    from cgat.functional import IMPL_DIRECT, conv2d_nhwc

    n, h, w, cin, cout, k, pad, act = case
    torch.manual_seed(23)
    x = (torch.rand(n, h, w, cin) - 0.5).to(dtype).float()
    wt = (torch.rand(cout, k, k, cin) - 0.5).to(dtype).float()
    b = torch.rand(cout) - 0.5
    xr, wr, br = (t.clone().requires_grad_() for t in (x, wt, b))
    pt, pl, pb, pr = pad
    yr = _act(F.conv2d(F.pad(xr.permute(0, 3, 1, 2), (pl, pr, pt, pb)), wr.permute(0, 3, 1, 2), br), act).permute(0, 2, 3, 1)
    g = (torch.rand_like(yr) - 0.5).to(dtype).float()
    yr.backward(g)
    xo = x.to(DEV, dtype).requires_grad_()
    wo = wt.to(DEV).requires_grad_()
    bo = b.to(DEV).requires_grad_()
    _lib.profile_start()
    yo = conv2d_nhwc(xo, wo, bo, stride=1, pad=pad, act=act, impl=IMPL_DIRECT)
    yo.backward(g.to(DEV, dtype))
    torch.cuda.synchronize()
    _lib.profile_stop()
    lo = dtype == torch.bfloat16
    close(yo, yr.detach(), rtol=1e-2 if lo else 1e-4, atol=1e-2 if lo else 1e-5, msg="y")
    close(xo.grad, xr.grad, rtol=1e-2 if lo else 1e-4, atol=2e-2 if lo else 1e-5, msg="dx")


def test_conv_small_is_the_kernel_that_serves_these_shapes():
    """The C ABI routes impl 0 for these shapes to conv_small_kernel: bit-identical results whether the caller passes the
    tensors as fprop of (cin=4 -> cout=4) or the library is asked again (determinism), and the generic implicit-GEMM path
    (cin = 12: not served) agrees with it on a zero-padded copy of the same problem."""
    from cgat.functional import IMPL_DIRECT, conv2d_nhwc

    torch.manual_seed(5)
    x = torch.rand(2, 32, 32, 4, device=DEV) - 0.5
    wt = torch.rand(4, 4, 4, 4, device=DEV) - 0.5
    y1 = conv2d_nhwc(x, wt, None, stride=1, pad=(1, 1, 2, 2), impl=IMPL_DIRECT)
    y2 = conv2d_nhwc(x, wt, None, stride=1, pad=(1, 1, 2, 2), impl=IMPL_DIRECT)
    assert torch.equal(y1, y2)
    xp = torch.cat([x, torch.zeros(2, 32, 32, 8, device=DEV)], -1)      # 12 input channels: conv_gemm_kernel
    wp = torch.cat([wt, torch.zeros(4, 4, 4, 8, device=DEV)], -1)
    y3 = conv2d_nhwc(xp, wp, None, stride=1, pad=(1, 1, 2, 2), impl=IMPL_DIRECT)
    close(y1, y3, rtol=1e-5, atol=1e-6, msg="conv_small vs conv_gemm")
