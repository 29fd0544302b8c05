import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from cgat import _lib
from cgat.functional import AttnConfig, graph_attention
from oracle import spec
torch.manual_seed(0)
nodes, ci, co, heads = 4, 6, 6, 1
P = 512
for proj, scale in [("pre", 1.0), ("pre", 3.0), ("lin", 1.0), ("lin", 3.0)]:
    a = (torch.rand(heads, 2 * co) - 0.5) * 2
    B = torch.rand(heads, nodes, nodes) * 0.3
    W = (torch.rand(heads, ci, co) - 0.5) * 2
    if proj == "pre":
        inp = ((torch.rand(1, P, nodes, co) - 0.3) * scale).bfloat16().float()   # Wh[n,p,node,c]
        Wh = inp.clone().requires_grad_()
    else:
        inp = (torch.rand(1, P, nodes, ci) * scale).bfloat16().float()
        X = inp.clone().requires_grad_()
        Wh = X @ W[0]
    out_r = spec.attention_core(Wh, a[0], spec.adjacency_norm(B[0]))
    g = (torch.rand_like(out_r) - 0.5).bfloat16().float()
    out_r.backward(g)
    din_r = (Wh if proj == "pre" else X).grad
    cfg = AttnConfig(nodes=nodes, ci=ci, co=co, heads=heads, layout=_lib.LAYOUT_TEMPORAL, proj=_lib.PROJ_PRE if proj == "pre" else _lib.PROJ_LINEAR,
                     merge=_lib.MERGE_MEAN, pix_per_sample=P)
    xi = inp.reshape(P, -1).cuda().bfloat16().requires_grad_()
    out = graph_attention(xi, None if proj == "pre" else W.cuda(), a.cuda(), B.cuda(), None, cfg)
    out.backward(g.reshape(P, -1).cuda().bfloat16())
    eo = (out.float().cpu().reshape(out_r.shape) - out_r.detach()).abs().max().item()
    ed = (xi.grad.float().cpu().reshape(din_r.shape) - din_r).abs()
    print(proj, scale, "out err %.4f/%.2f  din err max %.4f mean %.5f / max %.2f" % (eo, out_r.abs().max(), ed.max(), ed.mean(), din_r.abs().max()))
