"""GPU parity of cgat_s2d_pad (csrc/layout_kernels.cu) against the PyTorch formulation it replaces: F.pad by one pixel +
view / permute / reshape into 2x2 pixel blocks (the regrouping behind the DCGAN discriminators' stride-2 convs,
dcgan/model.py:150-165), forward and backward, bit-exact (a permutation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _torch_s2d(x):
    N, H, W, C = x.shape
    hs, ws = H // 2 + 1, W // 2 + 1
    xp = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))
    return xp.view(N, hs, 2, ws, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, hs, ws, 4 * C)


@pytest.mark.parametrize("shape", [(3, 8, 6, 4), (2, 64, 64, 4), (5, 16, 16, 64), (2, 4, 4, 24), (1, 2, 2, 3)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_s2d_pad_forward_and_backward_are_the_torch_permutation(shape, dtype):
    from cgat.conv_layers import _S2DPad

    torch.manual_seed(sum(shape))
    x = torch.randn(*shape, device=DEV).to(dtype)
    xr = x.clone().requires_grad_()
    xo = x.clone().requires_grad_()
    yr = _torch_s2d(xr)
    yo = _S2DPad.apply(xo)
    assert yo.shape == yr.shape and torch.equal(yo, yr)
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g)
    assert torch.equal(xo.grad, xr.grad)
