"""Developer aid: print the pipeline timeline (clock64 deltas) of CTA 0 of the tcgen05 fprop kernel."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "extended-gan_b200")):
    sys.path.insert(0, p)
import torch
from cgat import _lib
from cgat.functional import IMPL_TC, conv2d_nhwc

L = _lib.lib()
L.cgat_debug_timeline.argtypes = [ctypes.c_void_p]
L.cgat_debug_timeline.restype = None
x = torch.rand(64, 64, 64, 24, device="cuda").bfloat16()
w = (torch.rand(72, 3, 3, 24, device="cuda") - 0.5)
b = torch.rand(72, device="cuda")
for _ in range(2):
    conv2d_nhwc(x, w, b, pad=(1, 1, 1, 1), impl=IMPL_TC)
torch.cuda.synchronize()
buf = torch.zeros(16 * 9, dtype=torch.int64, device="cuda")
L.cgat_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
which = sys.argv[1] if len(sys.argv) > 1 else "fprop"
if which == "fprop":
    conv2d_nhwc(x, w, b, pad=(1, 1, 1, 1), impl=IMPL_TC)
else:
    xg = x.clone().requires_grad_(False)
    wg = w.clone().requires_grad_()
    y = conv2d_nhwc(xg, wg, b, pad=(1, 1, 1, 1), impl=IMPL_TC)
    torch.cuda.synchronize()
    buf.zero_()
    y.backward(torch.rand_like(y))
torch.cuda.synchronize()
L.cgat_debug_timeline(None)
t = buf.cpu().view(16, 9)
t0 = int(t[0, 0])
names = ["P:top", "P:empty", "P:rows", "P:relay", "M:tempty", "M:full", "M:issued", "E:tfull", "E:done"]
print("tile " + " ".join(f"{n:>9s}" for n in names))
for i in range(16):
    if int(t[i, 0]) == 0:
        continue
    print(f"{i:4d} " + " ".join(f"{(int(v) - t0) if int(v) else 0:9d}" for v in t[i]))
