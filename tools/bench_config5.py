"""BASELINE config 5: one DCGAN adversarial step (dcgan/train.py:97-160) on the conv kernels at the reference's size
(nc = 4, ndf = 64, 64 x 64 frames), fp32 and bf16 activations -- wall time per step and the conv kernels' share
(builder's tool, not the bench contract)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat import _lib
from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator
from dcgan.train import adversarial_step, default_criterion, make_optimizers

dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for dtype in (torch.float32, torch.bfloat16):
    torch.manual_seed(369)
    params = {"nc": 4, "ndf": 64}
    nets = [Generator(params).to(dev), FrameDiscriminator(params).to(dev), TemporalDiscriminator(params).to(dev)]
    oG, oFD, oTD = make_optimizers(*nets)
    x = torch.rand(N, 4, 64, 64, device=dev).to(dtype)
    y = torch.rand(N, 4, 64, 64, device=dev).to(dtype)
    crit = default_criterion()

    def step():
        return adversarial_step(netG=nets[0], netFD=nets[1], netTD=nets[2], optimizerG=oG, optimizerFD=oFD, optimizerTD=oTD,
                                criterion=crit, x=x, y=y)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    _lib.profile_start(); step(); torch.cuda.synchronize(); prof = _lib.profile_stop()
    # the same step replayed from a CUDA graph (host out of the loop)
    from dcgan.train import GraphedAdversarialStep
    torch.manual_seed(369)
    nets = [Generator(params).to(dev), FrameDiscriminator(params).to(dev), TemporalDiscriminator(params).to(dev)]
    oG, oFD, oTD = make_optimizers(*nets, capturable=True)
    gstep = GraphedAdversarialStep(netG=nets[0], netFD=nets[1], netTD=nets[2], optimizerG=oG, optimizerFD=oFD,
                                   optimizerTD=oTD, criterion=crit, x=x, y=y)
    for _ in range(2):
        gstep(x, y)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        out = gstep(x, y)
    b.record()
    torch.cuda.synchronize()
    gms = a.elapsed_time(b) / 10
    print(f"  graph replay: {gms:.2f} ms/step ({N / gms * 1e3:.0f} samples/s), errG {out[2].item():.4f}")
    print(f"DCGAN step N={N} {dtype}: {ms:.1f} ms/step ({N / ms * 1e3:.0f} samples/s); conv kernels (ms):",
          {k: round(c * t, 1) for k, (c, t) in prof.items()})
