"""Oracle restatement vs the live, unmodified reference on fresh random inputs (container only:
skipped where /root/reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_loader, spec
from util import close

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.mark.parametrize("seed", [1, 2])
def test_gat2d_random(seed):
    bm = ref_loader.baseline_model()
    torch.manual_seed(seed)
    N, P, T, V = 2, 50, 4, 5
    lay = bm.GraphAttentionLayer2D(T, 3, V, 0.2)
    with torch.no_grad():
        lay.B.add_(torch.rand(V, V))
    h = torch.rand(N, P, T, V, requires_grad=True)
    with ref_loader.cpu_shim():
        out = lay(h)
    g = torch.rand_like(out)
    out.backward(g)
    h2 = h.detach().clone().requires_grad_()
    W, a, B = (p.detach().clone().requires_grad_() for p in (lay.W, lay.a, lay.B))
    o2 = spec.gat2d_layer(h2, W, a, B)
    o2.backward(g)
    close(o2, out.detach())
    close(h2.grad, h.grad, atol=1e-5)
    close(W.grad, lay.W.grad, atol=1e-4)
    close(a.grad, lay.a.grad, atol=1e-4)
    close(B.grad, lay.B.grad, atol=1e-4)


def test_multihead2d_concat_axis():
    bm = ref_loader.baseline_model()
    torch.manual_seed(3)
    N, P, T, V = 2, 30, 4, 6
    mh = bm.GATMultiHead2D(T, T, V, 0.2, 3)
    h = torch.rand(N, P, T, V)
    with ref_loader.cpu_shim():
        out = mh(h)
    outs = [spec.gat2d_layer(h, m.W, m.a, m.B) for m in mh.attentions]
    close(torch.cat(outs, dim=2), out.detach())
