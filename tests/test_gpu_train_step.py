"""GPU parity of the train step (convolutional_gat/train.py:129-133, :212): forward, MSE - 0.0005*mean loss, backward,
Adam(weight_decay=0.01) -- cgat.train_step.TrainStep (CUDA graph; fused forward+loss+backward kernel where served)
against the CPU oracle (oracle/spec.py model + torch.optim.Adam)."""
import pytest
import torch

from oracle import spec
from util import close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _models(attention_type, mapping, seed=31):
    from convolutional_gat.GAT3D.GATMultistream import Model

    torch.manual_seed(seed)
    ours = Model(image_width=24, image_height=32, n_vertices=6, attention_type=attention_type, mapping_type=mapping)
    ref = spec.SpecGATMultiHead3D(4, 4, 0.2, 3, type_=attention_type, mapping_type=mapping, n_vertices=6)
    # the wrapper's hidden layer is the whole forward (model.py:44-47); give the oracle the same parameters
    ref.load_state_dict(ours.net.hidden_layer.state_dict())
    return ours.to(DEV), ref


@pytest.mark.parametrize("attention_type,mapping,fuse", [("temporal", "conv", True), ("temporal", "conv", False),
                                                         ("spatial", "conv", True), ("temporal", "linear", True)])
def test_train_steps_match_oracle(attention_type, mapping, fuse):
    from cgat.train_step import TrainStep

    ours, ref = _models(attention_type, mapping)
    torch.manual_seed(7)
    x = torch.rand(4, 32, 24, 4, 6).bfloat16()
    y = torch.rand(4, 32, 24, 4, 6).bfloat16()
    ts = TrainStep(ours, x.to(DEV), y.to(DEV), lr=1e-3, use_graph=True, fuse_loss=fuse)
    assert (ts.fused_stream is not None) == (fuse and mapping == "conv")
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=0.01)  # train.py:212
    for it in range(3):
        opt.zero_grad()
        out = ref(x.float())
        loss_r = spec.train_loss(out, y.float())  # train.py:131
        loss_r.backward()
        if it == 0:
            grads_r = {k: p.grad.clone() for k, p in ref.named_parameters()}
        opt.step()
        loss_o = ts.step(x.to(DEV), y.to(DEV))
        if it == 0:
            torch.cuda.synchronize()
            for k, p in ours.net.hidden_layer.named_parameters():
                g = grads_r[k]
                close(p.grad, g, rtol=2e-2, atol=3e-2 * max(1e-7, g.abs().max().item()), msg=f"step-0 d{k}")
        close(loss_o[0], loss_r.detach(), rtol=2e-2, atol=1e-4, msg=f"loss at step {it}")
    # Adam's first steps move every parameter by ~lr whatever the gradient's size: compare the parameters themselves
    pr = dict(ref.named_parameters())
    for k, p in ours.net.hidden_layer.named_parameters():
        close(p, pr[k].detach(), rtol=2e-2, atol=2.5e-3, msg=f"param {k} after 3 steps")
    # the unused output layer (model.py:44-47) gets no gradient and is not touched by the optimiser
    assert all(p.grad is None or not p.grad.any() for p in ours.net.output_layer.parameters())


@pytest.mark.parametrize("N,H,W", [(8, 64, 64), (14, 64, 64), (5, 50, 45), (64, 64, 64)])
def test_fused_loss_kernel_equals_unfused_step(N, H, W):
    """Same batch through cgat_layer_train (tile pairs in packed half2) and through layer_fwd + loss + layer_bwd
    (fp32 math).  The sizes give CTAs with 1-2, 3-4 and 13-14 tiles (odd counts leave a pair half empty) and ragged
    tiles in both directions."""
    from cgat.train_step import TrainStep

    res = {}
    for fuse in (True, False):
        ours, _ = _models("temporal", "conv", seed=33)
        torch.manual_seed(9)
        x = torch.rand(N, H, W, 4, 6, device=DEV).bfloat16()
        y = torch.rand(N, H, W, 4, 6, device=DEV).bfloat16()
        ts = TrainStep(ours, x, y, use_graph=False, fuse_loss=fuse)
        ts._fwd_bwd()
        torch.cuda.synchronize()
        res[fuse] = (ts.loss.clone(), ts.flat_grad.clone(), ts.mse.clone())
    close(res[True][0], res[False][0], rtol=2e-3, atol=1e-5, msg="loss")
    close(res[True][2], res[False][2], rtol=2e-3, atol=1e-5, msg="mse")
    close(res[True][1], res[False][1], rtol=2e-2, atol=3e-2 * res[False][1].abs().max().item(), msg="flat gradient")


@pytest.mark.parametrize("attention_type,N,H,W", [("temporal", 5, 50, 45), ("spatial", 6, 40, 64)])
def test_train_step_accepts_records_and_loader_planar_x(attention_type, N, H, W):
    """cgat_layer_train reads x padded chunk-planar.  TrainStep.step(x, y) with record tensors (converted by
    cgat_records_to_planar) and a slot filled through the loader kernel's planar output must give the same step."""
    from cgat.train_step import TrainStep
    from convolutional_gat.data_loaders.kmni_data_loader import gather_windows

    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 255, (N + 7, 6, H, W), generator=g, dtype=torch.uint8).to(DEV)
    start = torch.arange(N, dtype=torch.int32, device=DEV)
    x, y = gather_windows(frames, start, dtype=torch.bfloat16)
    res = {}
    for via_loader in (False, True):
        ours, _ = _models(attention_type, "conv", seed=35)
        ts = TrainStep(ours, x, y, use_graph=False)
        assert ts.fused_stream is not None and ts.xp is not None
        if via_loader:
            ts.xp.zero_()
            gather_windows(frames, start, out=(ts.xp, ts.y), planar=True)
        ts._fwd_bwd()
        torch.cuda.synchronize()
        res[via_loader] = (ts.loss.clone(), ts.flat_grad.clone(), ts.mse.clone())
    close(res[True][0], res[False][0], rtol=1e-5, atol=1e-7, msg="loss")
    close(res[True][2], res[False][2], rtol=1e-5, atol=1e-7, msg="mse")
    close(res[True][1], res[False][1], rtol=1e-5, atol=1e-6 * res[False][1].abs().max().item(), msg="flat gradient")


def test_loss_mirror_writes_every_steps_loss_to_host_memory():
    """enable_loss_mirror: the fused step's last launch (cgat_stream_finish_mirror) writes each step's loss into a pinned host
    ring -- same values as the device-side loss read back step by step, graph replay and eager, warm-ups not counted."""
    from cgat.train_step import TrainStep

    ours, _ = _models("temporal", "conv")
    torch.manual_seed(11)
    x = torch.rand(4, 32, 24, 4, 6, device=DEV).bfloat16()
    y = torch.rand(4, 32, 24, 4, 6, device=DEV).bfloat16()
    ts = TrainStep(ours, x, y, lr=1e-3, use_graph=True)
    ring = ts.enable_loss_mirror(4)
    assert ring.is_pinned() and ring.numel() == 4
    seen = []
    for it in range(6):  # wraps around the ring of 4
        seen.append(float(ts.run()[0]))  # (.item() synchronises: the posted host write of this step has landed)
        torch.cuda.synchronize()
        assert ts.loss_of_step(it) == seen[-1], (it, ts.loss_of_step(it), seen[-1])
    assert seen[0] != seen[-1]  # the parameters moved
    assert int(ts._mirror[1]) == 6
