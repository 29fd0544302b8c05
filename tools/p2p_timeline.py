"""Where the data-parallel step's extra time goes: per-step globaltimer stamps of the peer-memory exchange kernel
(start, pushed, all sources arrived for element 0, end) on every rank, and the gap to the next step's kernel
(= the step's graph).  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_timeline.py
(builder's tool)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    from cgat import _lib
    from cgat.train_step import TrainStep
    from convolutional_gat.GAT3D.GATMultistream import Model

    torch.manual_seed(369)
    model = Model(image_width=64, image_height=64, n_vertices=6, attention_type="temporal", mapping_type="conv").to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(64, 64, 64, 4, 6, generator=g).bfloat16().to(dev)
    y = torch.rand(64, 64, 64, 4, 6, generator=g).bfloat16().to(dev)
    ts = TrainStep(model, x, y, lr=1e-3)
    ts.sync_params()
    dbg = torch.zeros(4096 * 4, dtype=torch.int64, device=dev)
    _lib.lib().cgat_debug_timeline(ctypes.c_void_p(dbg.data_ptr()))  # (before the exchange kernel is captured into the step's graph)
    assert ts.enable_p2p_exchange()
    K = 200
    for _ in range(K):
        ts.run()
    torch.cuda.synchronize()
    _lib.lib().cgat_debug_timeline(None)
    t = dbg.view(4096, 4)[50:K].cpu().double() / 1e3  # us; skip warm-up
    push, wait, tail = (t[:, 1] - t[:, 0]), (t[:, 2] - t[:, 1]), (t[:, 3] - t[:, 2])
    gap = t[1:, 0] - t[:-1, 3]
    step = t[1:, 0] - t[:-1, 0]
    med = lambda v: float(v.median())
    msg = (f"rank {rank}/{world}: step {med(step):.1f} us = exchange kernel {med(t[:, 3] - t[:, 0]):.1f} (push {med(push):.1f}, "
           f"wait {med(wait):.1f}, adam {med(tail):.1f}) + rest of the step {med(gap):.1f}; wait p90 {float(wait.quantile(0.9)):.1f}")
    out = [None] * world
    dist.all_gather_object(out, msg)
    if rank == 0:
        print("\n".join(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
