"""One UnetModel [2,128,128,4,8] forward + backward (BASELINE config 4, bf16 activations over fp32 parameters, batched
vertices) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_config4_unet.py
(builder's tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from convolutional_gat.unet_model import UnetModel

dev = "cuda"
torch.manual_seed(369)
dt = torch.bfloat16 if len(sys.argv) < 2 or sys.argv[1] != "fp32" else torch.float32
m = UnetModel(image_width=128, image_height=128, n_vertices=8, attention_type="unet").to(dev)
x = torch.rand(2, 128, 128, 4, 8, device=dev).to(dt).requires_grad_()


def step():
    out = m(x)
    out.backward(torch.ones_like(out))


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
