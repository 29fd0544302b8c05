"""One DCGAN adversarial step (BASELINE config 5, bf16, N = 64) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_config5.py
(launch list of every kernel of the step; builder's tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from dcgan.model import FrameDiscriminator, Generator, TemporalDiscriminator
from dcgan.train import adversarial_step, default_criterion, make_optimizers

dev = "cuda"
N = 64
torch.manual_seed(369)
params = {"nc": 4, "ndf": 64}
nets = [Generator(params).to(dev), FrameDiscriminator(params).to(dev), TemporalDiscriminator(params).to(dev)]
oG, oFD, oTD = make_optimizers(*nets)
x = torch.rand(N, 4, 64, 64, device=dev).bfloat16()
y = torch.rand(N, 4, 64, 64, device=dev).bfloat16()
crit = default_criterion()
step = lambda: adversarial_step(netG=nets[0], netFD=nets[1], netTD=nets[2], optimizerG=oG, optimizerFD=oFD,
                                optimizerTD=oTD, criterion=crit, x=x, y=y)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
