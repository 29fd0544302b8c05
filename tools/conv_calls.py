"""Per-call timing of the conv entry points during one DCGAN adversarial step (dcgan/train.py:97-160) or, with a third
argument "unet", one forward+backward of UnetModel (unet_model.py:22-29) at 128x128: which conv shapes take which kernel
class and what each costs (builder's tool).  Usage: conv_calls.py N bf16|fp32 [unet]"""
import ctypes, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat import _lib
from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator
from dcgan.train import adversarial_step, default_criterion, make_optimizers

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dtype = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
torch.manual_seed(369)
UNET = len(sys.argv) > 3 and sys.argv[3] == "unet"
params = {"nc": 4, "ndf": 64}
nets = [Generator(params).to(dev), FrameDiscriminator(params).to(dev), TemporalDiscriminator(params).to(dev)]
oG, oFD, oTD = make_optimizers(*nets)
x = torch.rand(N, 4, 64, 64, device=dev).to(dtype)
y = torch.rand(N, 4, 64, 64, device=dev).to(dtype)
crit = default_criterion()
step = lambda: adversarial_step(netG=nets[0], netFD=nets[1], netTD=nets[2], optimizerG=oG, optimizerFD=oFD,
                                optimizerTD=oTD, criterion=crit, x=x, y=y)
if UNET:
    from convolutional_gat.unet_model import UnetModel
    um = UnetModel(image_width=128, image_height=128, n_vertices=8, attention_type="unet").to(dev).to(dtype)
    ux = torch.rand(N, 128, 128, 4, 8, device=dev).to(dtype).requires_grad_()

    def step():
        out = um(ux)
        out.backward(torch.ones_like(out))
for _ in range(2):
    step()
torch.cuda.synchronize()
records = []
orig = _lib.call


def traced(name, *args, launches=1):
    if name.startswith("cgat_conv2d"):
        d = args[0]._obj
        impl = args[5] if name.endswith("fprop") else (args[4] if name.endswith("dgrad") else args[5])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); orig(name, *args, launches=launches); b.record()
        records.append((name[12:], (d.n, d.h, d.w, d.cin, d.cout, d.kh, d.stride, d.groups), impl, a, b))
    else:
        orig(name, *args, launches=launches)


_lib.call = traced
import cgat.functional as Fm
Fm._lib.call = traced
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record()
torch.cuda.synchronize()
print(f"step {a.elapsed_time(b):.2f} ms (traced, serialised by events)")
agg = collections.OrderedDict()
for which, shp, impl, e0, e1 in records:
    k = (which, shp, impl)
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + e0.elapsed_time(e1))
tot = 0
for (which, shp, impl), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tot += t
    print(f"{which:6s} n,h,w,cin,cout,k,s,g={shp} impl={impl} calls={c} total={t:.3f} ms")
print(f"conv total {tot:.2f} ms")
