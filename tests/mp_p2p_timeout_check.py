"""Multi-process GPU check of the exchange kernel's TIME-OUT path (run under torchrun, 2 ranks): rank 0 takes a step that
rank 1 never takes.  Rank 0's cluster must give up after ~3 s, leave parameters, Adam moments and the device step counter
untouched (all-or-nothing across the cluster's CTAs), set the time-out marker, and TrainStep must turn the marker into a
RuntimeError at its next poll."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "extended-gan_b200")]
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cgat.train_step import TrainStep
    from convolutional_gat.GAT3D.GATMultistream import Model

    torch.manual_seed(369)
    model = Model(image_width=24, image_height=32, n_vertices=6, attention_type="temporal", mapping_type="conv").to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(4, 32, 24, 4, 6, generator=g).bfloat16().to(dev)
    y = torch.rand(4, 32, 24, 4, 6, generator=g).bfloat16().to(dev)
    ts = TrainStep(model, x, y, lr=1e-3)
    ts.sync_params()
    assert ts.enable_p2p_exchange(), getattr(ts.flat, "p2p_error", "p2p setup returned False")
    ts.step(x, y)  # one good step on both ranks
    torch.cuda.synchronize()
    dist.barrier()
    assert not ts.flat.p2p_timed_out()
    before = (ts.flat_param.clone(), ts.exp_avg.clone(), ts.exp_avg_sq.clone(), int(ts._step_dev))
    assert before[3] == 1
    if rank == 0:
        t0 = time.time()
        ts.step(x, y)  # the peer never takes this step
        torch.cuda.synchronize()
        waited = time.time() - t0
        assert ts.flat.p2p_timed_out(), "the exchange kernel did not record its time-out"
        assert 1.0 < waited < 30.0, waited
        assert torch.equal(ts.flat_param, before[0]) and torch.equal(ts.exp_avg, before[1]) and torch.equal(ts.exp_avg_sq, before[2])
        assert int(ts._step_dev) == 1, "the device step counter advanced on a timed-out step"
        raised = False
        try:
            for _ in range(2 * ts.P2P_POLL_EVERY + 2):  # the poll is asynchronous: refreshed every P2P_POLL_EVERY steps
                ts._exchange_and_update()
                torch.cuda.synchronize()
        except RuntimeError as e:
            raised = "timed out" in str(e)
        assert raised, "TrainStep did not raise on the time-out marker"
        print(f"p2p time-out OK: gave up after {waited:.1f} s, nothing updated, RuntimeError raised")
    else:
        time.sleep(8.0)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
