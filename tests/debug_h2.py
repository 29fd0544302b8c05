import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from test_gpu_attention import _pair
for type_, merge, mapping in [("temporal", "mean", "linear"), ("spatial", "concat", "linear"), ("temporal", "concat", "linear"), ("temporal", "mean", "conv")]:
    ours, ref = _pair(type_, mapping, 3, merge, "neighbour", False, seed=13)
    x = torch.rand(2, 16, 16, 4, 6).bfloat16().float()
    xr = x.clone().requires_grad_()
    out_r = ref(xr); g = torch.rand_like(out_r) - 0.5; out_r.backward(g)
    res = {}
    for env in ("0", "1"):

        xo = x.cuda().bfloat16().requires_grad_()
        out_o = ours(xo); out_o.backward(g.cuda().bfloat16())
        e_out = (out_o.float().cpu() - out_r.detach()).abs().max().item()
        e_dx = (xo.grad.float().cpu() - xr.grad).abs().max().item()
        res[env] = (e_out, e_dx)
        break
    print(type_, merge, mapping, "out err %.4f (max %.2f)  dx err %.4f (max %.2f, rms %.3f)" % (res["0"][0], out_r.abs().max(), res["0"][1], xr.grad.abs().max(), xr.grad.pow(2).mean().sqrt()))
