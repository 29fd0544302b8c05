"""The bench shape's loader gather (64 windows of 64x64x4x6 from 71 uint8 frames, chunk-planar x + record y, bf16) a few
times, for ncu (builder's tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat.functional import planar_zeros
from convolutional_gat.data_loaders.kmni_data_loader import gather_windows

dev = "cuda"
g = torch.Generator().manual_seed(369)
frames = torch.randint(0, 255, (71, 6, 64, 64), generator=g, dtype=torch.uint8).to(dev)
start = torch.arange(64, dtype=torch.int32, device=dev)
xp = planar_zeros((64, 64, 64, 4, 6), dev)
y = torch.empty(64, 64, 64, 4, 6, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    gather_windows(frames, start, steps=4, out=(xp, y), planar=True)
torch.cuda.synchronize()
