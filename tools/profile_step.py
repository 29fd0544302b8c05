"""Run a few EAGER conv-GAT train steps (BASELINE config 2 shapes) so ncu can attribute kernels by name.

    python tools/profile_step.py [--steps 3] [--mapping conv] [--type temporal] [--batch 64]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "extended-gan_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from bench import SEED, SHAPE, synth_batch  # noqa: E402
from cgat.train_step import TrainStep  # noqa: E402
from convolutional_gat.GAT3D.GATMultistream import Model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--type", default="temporal")
ap.add_argument("--mapping", default="conv")
args = ap.parse_args()

H, W, T, V = SHAPE
dev = torch.device("cuda", 0)
torch.manual_seed(SEED)
model = Model(image_width=W, image_height=H, n_vertices=V, attention_type=args.type, mapping_type=args.mapping).to(dev)
x, y = synth_batch(args.batch, torch.bfloat16)
ts = TrainStep(model, x.to(dev), y.to(dev), use_graph=False)
for _ in range(args.steps):
    loss = ts.run()
torch.cuda.synchronize()
print("loss", float(loss))
