"""Where the end-to-end step's extra time over the device-resident step goes (config 2, one GPU): the pipelined graph
(train step || H2D copies + loader gather of the next batch) with pieces removed (builder's tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
from cgat.train_step import TrainStep
from convolutional_gat.GAT3D.GATMultistream import Model
from convolutional_gat.data_loaders.kmni_data_loader import gather_windows

dev = torch.device("cuda")
torch.manual_seed(369)
B, H, W, T, V = 64, 64, 64, 4, 6
model = Model(image_width=W, image_height=H, n_vertices=V, attention_type="temporal", mapping_type="conv").to(dev)
x = torch.rand(B, H, W, T, V, device=dev).bfloat16(); y = torch.rand(B, H, W, T, V, device=dev).bfloat16()
ts = TrainStep(model, x, y, lr=1e-3, use_graph=True)
frames_h = torch.randint(0, 255, (B + 7, V, H, W), dtype=torch.uint8).pin_memory()
start_h = torch.arange(B, dtype=torch.int32).pin_memory()
ts.enable_prefetch(2)
for s in ts._slots[:2]:
    s["frames"] = torch.empty(frames_h.shape, dtype=torch.uint8, device=dev)
    s["start"] = torch.empty(start_h.shape, dtype=torch.int32, device=dev)
    s["frames"].copy_(frames_h); s["start"].copy_(start_h)
loss_host = torch.empty(4096, dtype=torch.float32).pin_memory()

def build(copies, gather, gather_after_layer=False):
    side = torch.cuda.Stream()
    graphs = []
    keep = (ts.x, ts.y, ts.xp, ts.graph)
    torch.cuda.synchronize()
    for k in range(2):
        cur_slot, other = ts._slots[k], ts._slots[1 - k]
        ts.x, ts.y, ts.xp = cur_slot["x"], cur_slot["y"], cur_slot["xp"]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if copies:
                    other["frames"].copy_(frames_h, non_blocking=True)
                    other["start"].copy_(start_h, non_blocking=True)
                if gather:
                    gather_windows(other["frames"], other["start"], crop=H, steps=T, out=(other["xp"], other["y"]), planar=True)
            ts._step_launches()
            cur.wait_stream(side)
        graphs.append(g)
    ts.x, ts.y, ts.xp, ts.graph = keep
    return graphs

def timeit(fn, K=400):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K * 1e3

print(f"plain step (graph replay)                 {timeit(lambda i: ts.run()):7.1f} us")
def with_loss(i):
    ts.run(); loss_host[i:i + 1].copy_(ts.loss, non_blocking=True)
print(f"plain step + 4-byte D2H loss copy         {timeit(with_loss):7.1f} us")
for name, c, g in (("copies + gather (the e2e graph)", 1, 1), ("gather only", 0, 1), ("copies only", 1, 0), ("neither", 0, 0)):
    gs = build(c, g)
    def run(i, gs=gs):
        gs[i & 1].replay(); ts._exchange_and_update()
    def run_loss(i, gs=gs):
        gs[i & 1].replay(); ts._exchange_and_update(); loss_host[i:i + 1].copy_(ts.loss, non_blocking=True)
    print(f"pipelined graph, {name:32s} {timeit(run):7.1f} us   + loss copy {timeit(run_loss):7.1f} us")
