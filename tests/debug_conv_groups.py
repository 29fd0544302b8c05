import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch, torch.nn as nn
from cgat.conv_layers import Conv2d
torch.manual_seed(0)
cases = [dict(cin=8, cout=16, k=3, pad=1, groups=8), dict(cin=2, cout=1, k=7, pad=3, groups=1, bias=False),
         dict(cin=64, cout=32, k=1, pad=0, groups=1), dict(cin=4, cout=8, k=3, pad=1, groups=4),
         dict(cin=1024, cout=2048, k=3, pad=1, groups=1024), dict(cin=2048, cout=512, k=1, pad=0, groups=1)]
for c in cases:
    ref = nn.Conv2d(c["cin"], c["cout"], c["k"], padding=c["pad"], groups=c["groups"], bias=c.get("bias", True))
    ours = Conv2d(c["cin"], c["cout"], c["k"], padding=c["pad"], groups=c["groups"], bias=c.get("bias", True))
    ours.load_state_dict(ref.state_dict())
    ours = ours.cuda()
    hw = 4 if c["cin"] >= 1024 else 12
    x = torch.rand(2, c["cin"], hw, hw)
    xr = x.clone().requires_grad_(); xo = x.cuda().requires_grad_()
    yr = ref(xr); yo = ours(xo)
    g = torch.rand_like(yr)
    yr.backward(g); yo.backward(g.cuda())
    print(c, "y %.2e dx %.2e dw %.2e" % ((yo.cpu() - yr).abs().max(), (xo.grad.cpu() - xr.grad).abs().max(),
          (ours.weight.grad.cpu() - ref.weight.grad).abs().max()), "db %.2e" % ((ours.bias.grad.cpu() - ref.bias.grad).abs().max()) if ref.bias is not None else "")
