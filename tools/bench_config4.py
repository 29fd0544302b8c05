"""BASELINE config 4 (stress): UnetModel encoder and the many-node GAT layer at 128 x 128 -- fwd+bwd timings and the
kernels that dominate (CUDA events; builder's tool, not the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat import _lib
from cgat.layers import GATMultiHead3D
from convolutional_gat.unet_model import UnetModel

dev = "cuda"
torch.manual_seed(369)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for V in (32, 64):
    layer = GATMultiHead3D(4, 4, 0.2, 3, type_="spatial", mapping_type="linear", n_vertices=V).to(dev)
    x = torch.rand(4, 128, 128, 4, V, device=dev).bfloat16().requires_grad_()

    def step():
        out = layer(x)
        out.backward(torch.ones_like(out))

    ms = timeit(step)
    _lib.profile_start(); step(); torch.cuda.synchronize(); prof = _lib.profile_stop()
    print(f"GAT layer V={V} linear, x[4,128,128,4,{V}] bf16: fwd+bwd {ms:.3f} ms ->", {k: round(v[1], 3) for k, v in prof.items()})

for V, N, dt in ((8, 2, torch.float32), (8, 2, torch.bfloat16), (8, 16, torch.bfloat16)):
    m = UnetModel(image_width=128, image_height=128, n_vertices=V, attention_type="unet").to(dev).to(dt)
    x = torch.rand(N, 128, 128, 4, V, device=dev).to(dt).requires_grad_()

    def step():
        out = m(x)
        out.backward(torch.ones_like(out))

    ms = timeit(step, 2)
    _lib.profile_start(); step(); torch.cuda.synchronize(); prof = _lib.profile_stop()
    tot = {}
    for k, (c, t) in prof.items():
        tot[k] = round(c * t, 2)
    print(f"UnetModel V={V}, x[{N},128,128,4,{V}] {dt}: fwd+bwd {ms:.1f} ms; our kernels (ms total):", tot)

# the same model the way train(**cfg) runs it: TrainStep captures forward + loss + backward in one CUDA graph, so the
# host's ~3 000 launches per step (per-vertex loop, BatchNorm / CBAM / pooling glue) drop out of the step time
from cgat.train_step import TrainStep
for V, N, dt in ((8, 2, torch.float32),):  # TrainStep keeps fp32 master parameters: the PyTorch glue layers need fp32 inputs
    torch.manual_seed(369)
    m = UnetModel(image_width=128, image_height=128, n_vertices=V, attention_type="unet").to(dev).to(dt)
    x = torch.rand(N, 128, 128, 4, V, device=dev).to(dt)
    y = torch.rand(N, 128, 128, 4, V, device=dev).to(dt)
    ts = TrainStep(m, x, y, lr=1e-3, use_graph=True)
    ms = timeit(lambda: ts.run(), 5)
    print(f"UnetModel V={V}, x[{N},128,128,4,{V}] {dt}: full train step from a CUDA graph {ms:.1f} ms ({N / ms * 1e3:.1f} samples/s)")
