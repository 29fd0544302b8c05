"""Developer aid (not collected by pytest): prints max errors of the tcgen05 conv per case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import torch
from test_gpu_conv_tc import TC_CASES, _ref_conv
from cgat.functional import IMPL_TC, IMPL_DIRECT, conv2d_nhwc

for case in TC_CASES:
    n, h, w, cin, cout, k, pad = case
    torch.manual_seed(7)
    x = (torch.rand(n, h, w, cin) - 0.5).bfloat16().float()
    wt = (torch.rand(cout, k, k, cin) - 0.5).bfloat16().float()
    b = torch.rand(cout) - 0.5
    xr, wr, br = (t.clone().requires_grad_() for t in (x, wt, b))
    yr = _ref_conv(xr, wr, br, pad)
    g = (torch.rand_like(yr) - 0.5).bfloat16().float()
    yr.backward(g)
    yr = yr.detach()
    try:
        xo = x.cuda().bfloat16().requires_grad_(); wo = wt.cuda().requires_grad_(); bo = b.cuda().requires_grad_()
        yy = conv2d_nhwc(xo, wo, bo, stride=1, pad=pad, impl=IMPL_TC)
        yy.backward(g.cuda().bfloat16())
        torch.cuda.synchronize()
        ew = (wo.grad.cpu() - wr.grad).abs()
        print(case, "dw err %.4f (max %.2f)  db err %.4f (max %.2f) dx err %.4f (max %.2f)" % (ew.max(), wr.grad.abs().max(), (bo.grad.cpu() - br.grad).abs().max(), br.grad.abs().max(), (xo.grad.float().cpu() - xr.grad).abs().max(), xr.grad.abs().max()))
        if ew.max() > 0.02 * wr.grad.abs().max():
            bad = ew > 0.02 * wr.grad.abs().max()
            print("   bad dw per cout", bad.sum(dim=(1, 2, 3)).tolist()[:20], "per tap", bad.sum(dim=(0, 3)).tolist(), "per cin", bad.sum(dim=(0, 1, 2)).tolist()[:32])
    except Exception as e:  # noqa
        print(case, "BWD EXC", repr(e)[:300])
    try:
        yo = conv2d_nhwc(x.cuda().bfloat16(), wt.cuda(), b.cuda(), stride=1, pad=pad, impl=IMPL_TC).float().cpu()
        yd = conv2d_nhwc(x.cuda().bfloat16(), wt.cuda(), b.cuda(), stride=1, pad=pad, impl=IMPL_DIRECT).float().cpu()
        torch.cuda.synchronize()
        err = (yo - yr).abs()
        print(case, "tc max err %.4f (ref max %.3f) direct err %.4f  nan=%d" % (err.max(), yr.abs().max(), (yd - yr).abs().max(), torch.isnan(yo).sum()))
        if err.max() > 0.05:
            bad = (err > 0.05).nonzero()
            print("   first bad idx", bad[:5].tolist(), "count", len(bad), "of", err.numel())
            print("   bad per channel", (err > 0.05).sum(dim=(0, 1, 2))[:16].tolist())
            print("   bad per (h)", (err > 0.05).sum(dim=(0, 2, 3)).tolist())
            print("   bad per (w)", (err > 0.05).sum(dim=(0, 1, 3)).tolist())
    except Exception as e:  # noqa
        print(case, "EXC", repr(e)[:300])
