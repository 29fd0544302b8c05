"""UnetModel [2,128,128,4,8] (BASELINE config 4) full train step from TrainStep's CUDA graph: fp32, and bf16 activations
over fp32 master parameters; per-kernel totals of one eager forward + backward (builder's tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat import _lib
from convolutional_gat.unet_model import UnetModel
from cgat.train_step import TrainStep
dev = "cuda"
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for dt in (torch.float32, torch.bfloat16):
    torch.manual_seed(369)
    m = UnetModel(image_width=128, image_height=128, n_vertices=8, attention_type="unet").to(dev)
    x = torch.rand(2, 128, 128, 4, 8, device=dev).to(dt).requires_grad_()
    def step():
        out = m(x); out.backward(torch.ones_like(out))
    ms = timeit(step, 3)
    _lib.profile_start(); step(); torch.cuda.synchronize(); prof = _lib.profile_stop()
    tot = {k: (c, round(c * t, 2)) for k, (c, t) in prof.items()}
    tot = dict(sorted(tot.items(), key=lambda kv: -kv[1][1]))
    print(f"{dt} activations: eager fwd+bwd {ms:.1f} ms; our kernels total {sum(v[1] for v in tot.values()):.1f} ms, launches {sum(v[0] for v in tot.values())}")
    for k, v in list(tot.items())[:12]: print("    ", k, v)
    xx = torch.rand(2, 128, 128, 4, 8, device=dev).to(dt); yy = torch.rand(2, 128, 128, 4, 8, device=dev).to(dt)
    ts = TrainStep(m, xx, yy, lr=1e-3, use_graph=True)
    ms = timeit(lambda: ts.run(), 10)
    print(f"{dt} activations, fp32 parameters: TrainStep graph {ms:.2f} ms/step ({2 / ms * 1e3:.1f} samples/s), loss {float(ts.run()):.5f}")
