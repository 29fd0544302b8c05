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
    yr = _ref_conv(x, wt, b, pad)
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
