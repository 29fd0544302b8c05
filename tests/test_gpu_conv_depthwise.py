"""GPU parity of the vectorised depthwise 3x3 kernels (csrc/conv_depthwise.cu; the SmaAt-UNet DepthwiseSeparableConv
behind unet_model.py:20) vs torch CPU conv2d: fprop, dgrad, wgrad and dbias, fp32 (rtol 1e-4) and bf16 (2e-2)."""
import pytest
import torch
import torch.nn as nn

from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # n, h, w, cin, multiplier
    (2, 13, 11, 8, 2),      # ragged image, two channel octets
    (2, 16, 16, 4, 2),      # first conv of the encoder: 4 -> 8, one octet
    (2, 16, 16, 64, 2),     # inc block
    (3, 9, 9, 128, 1),      # multiplier 1
    (2, 4, 4, 1024, 2),     # deepest block: 256 octets, 80 KB of shared-memory accumulators
    (1, 4, 4, 1536, 2),     # weights too large for shared memory: generic direct kernels
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
def test_depthwise_3x3(case, dtype):
    from cgat.conv_layers import Conv2d

    n, h, w, cin, m = case
    torch.manual_seed(13)
    ref = nn.Conv2d(cin, cin * m, 3, padding=1, groups=cin)
    ours = Conv2d(cin, cin * m, 3, padding=1, groups=cin).to(DEV)
    with torch.no_grad():
        ref.weight.copy_(ref.weight.to(dtype).float())
        ours.weight.copy_(ref.weight)
        ours.bias.copy_(ref.bias)
    x = (torch.rand(n, cin, h, w) - 0.5).to(dtype)
    xr = x.float().clone().requires_grad_()
    yr = ref(xr)
    g = (torch.rand_like(yr) - 0.5).to(dtype).float()
    yr.backward(g)
    xo = x.clone().to(DEV).requires_grad_()
    yo = ours(xo)
    yo.backward(g.to(DEV, dtype))
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    close(yo, yr.detach(), rtol=tol, atol=tol, msg="y")
    close(xo.grad, xr.grad, rtol=tol, atol=tol, msg="dx")
    gs = max(1.0, ref.weight.grad.abs().max().item())
    close(ours.weight.grad, ref.weight.grad, rtol=tol, atol=tol * gs, msg="dw")
    close(ours.bias.grad, ref.bias.grad, rtol=tol, atol=tol * max(1.0, ref.bias.grad.abs().max().item()), msg="db")
