"""The BENCHED configuration (BASELINE.json configs[1]: x, y [64,64,64,4,6] bf16, GATMultistream.Model(conv), 3 heads)
compared with the CPU oracle (oracle/spec.py, fp32) AT ITS OWN SIZE: loss, every parameter gradient and the parameters
after the Adam step of cgat_layer_train (reference loop convolutional_gat/train.py:129-133, optimizer :212), and
output / gradients of cgat_layer_fwd / cgat_layer_bwd.

Tolerances: the north-star's rtol 2e-2 (bf16) plus an absolute floor that is MEASURED per tensor on the oracle, not a
fraction of max|g|: the oracle is run a second time with the roundings ANY bf16 implementation of this layer has to
make -- conv weights, projected features Wh and their gradient d(Wh) rounded to bf16 (straight-through) -- and the
RMS deviation of each parameter gradient from the exact fp32 run is that tensor's rounding noise sigma.  An element
passes if  |ours - oracle| <= 2e-2 |oracle| + 4 sigma.  For the conv weights sigma is ~1e-3 of the gradient's scale;
for the cancelling sums d(a), d(B) it is what cancellation leaves of 2^-9 relative errors on 262,144 pixel terms.
"""
import pytest
import torch

from oracle import spec

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_ULP = 2.0 ** -8
N, H, W, T, V = 64, 64, 64, 4, 6
LR = 1e-3


def _models(attention_type, seed=369):
    from convolutional_gat.GAT3D.GATMultistream import Model

    torch.manual_seed(seed)
    ours = Model(image_width=W, image_height=H, n_vertices=V, attention_type=attention_type, mapping_type="conv")
    ref = spec.SpecGATMultiHead3D(4, 4, 0.2, 3, type_=attention_type, mapping_type="conv", n_vertices=V)
    ref.load_state_dict(ours.net.hidden_layer.state_dict())
    return ours.to(DEV), ref


def _batch(scale=1.0, seed=11):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(N, H, W, T, V, generator=g) * scale).bfloat16()
    y = (torch.rand(N, H, W, T, V, generator=g) * scale).bfloat16()
    return x, y


class _RoundBf16(torch.autograd.Function):
    """Rounds to bf16 in the forward AND rounds the incoming gradient to bf16 in the backward."""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _grads(ref, run, emulate_bf16):
    """Parameter gradients of ``run(ref)`` (a scalar); with ``emulate_bf16`` the conv weights, Wh and d(Wh) are bf16."""
    orig = spec.node_conv3x3
    if emulate_bf16:
        spec.node_conv3x3 = lambda feat, w, b, H, W: _RoundBf16.apply(orig(feat, _RoundBf16.apply(w), b, H, W))
    try:
        for p in ref.parameters():
            p.grad = None
        val = run(ref)
        val.backward()
        return float(val.detach()), {k: p.grad.clone() for k, p in ref.named_parameters()}
    finally:
        spec.node_conv3x3 = orig


def _oracle_step(ref, x, y):
    """fp32 loss, gradients, per-tensor bf16 rounding noise sigma and the parameters after one Adam step."""
    xf, yf = x.float(), y.float()
    run = lambda m: spec.train_loss(m(xf), yf)  # train.py:131
    loss, total = _grads(ref, run, False)
    _, emu = _grads(ref, run, True)
    sigma = {k: float((emu[k] - total[k]).pow(2).mean().sqrt()) for k in total}
    opt = torch.optim.Adam(ref.parameters(), lr=LR, weight_decay=0.01)  # train.py:212
    for k, p in ref.named_parameters():
        p.grad = total[k].clone()
    opt.step()
    return loss, total, sigma, {k: p.detach().clone() for k, p in ref.named_parameters()}


def _assert_grad(name, got, want, sigma):
    got, want = got.float().cpu(), want.float()
    tol = 2e-2 * want.abs() + 4.0 * sigma
    err = (got - want).abs()
    bad = err > tol
    assert not bad.any(), (f"d{name}: {int(bad.sum())}/{bad.numel()} elements outside rtol 2e-2 + 4 sigma_bf16; worst "
                           f"err {float(err[bad].max()):.3e} vs tol {float(tol[bad][err[bad].argmax()]):.3e}, "
                           f"|g| max {float(want.abs().max()):.3e}")


@pytest.mark.parametrize("attention_type", ["temporal", "spatial"])
def test_config2_train_kernel_vs_oracle_full_size(attention_type):
    from cgat.train_step import TrainStep

    ours, ref = _models(attention_type)
    p0 = {k: p.detach().clone() for k, p in ref.named_parameters()}
    x, y = _batch()
    loss_r, grads_r, abs_r, params_r = _oracle_step(ref, x, y)
    ts = TrainStep(ours, x.to(DEV), y.to(DEV), lr=LR, use_graph=True)
    assert ts.fused_stream is not None, "the benched model must take the cgat_layer_train path"
    # TrainStep's construction probes run forward/backward only; parameters are still the initial ones
    for k, p in ours.net.hidden_layer.named_parameters():
        assert torch.equal(p.detach().cpu(), p0[k]), k
    loss_o = ts.step(x.to(DEV), y.to(DEV))
    torch.cuda.synchronize()
    assert not ts.range_guard_fired, "U[0,1) inputs are well inside the packed-fp16 kernel's range: no fp32 re-run"
    # loss: a mean of 6.3 M non-negative terms, no cancellation -> far inside the bf16 bar
    assert abs(float(loss_o[0]) - loss_r) <= 2e-3 * abs(loss_r), (float(loss_o[0]), loss_r)
    for k, p in ours.net.hidden_layer.named_parameters():
        _assert_grad(k, p.grad, grads_r[k], abs_r[k])
    # parameters after Adam: the first step moves every element by lr * g'/(|g'| + eps) ~ lr * sign(g') with
    # g' = g + 0.01 p (L2-style decay, train.py:212).  Where the oracle's |g'| clears the gradient's rounding floor the
    # update is determined: it must agree to 5 % of one step.  Where it does not, the sign of g' is not determined
    # and the element may land 2*lr away.
    for k, p in ours.net.hidden_layer.named_parameters():
        got, want = p.detach().float().cpu(), params_r[k]
        floor = 4.0 * abs_r[k] + 2e-2 * grads_r[k].abs()
        decided = (grads_r[k] + 0.01 * p0[k]).abs() > 2.0 * floor
        tol = torch.where(decided, torch.full_like(want, 0.05 * LR + 1e-6), torch.full_like(want, 2.0 * LR * 1.01))
        err = (got - want).abs()
        assert (err <= tol).all(), (f"param {k} after Adam: worst err {float(err.max()):.3e}, "
                                    f"{int((err > tol).sum())} of {err.numel()} outside; decided {int(decided.sum())}")
        assert decided.float().mean() > 0.5, f"{k}: the floor must leave most elements determined ({decided.float().mean():.2f})"


@pytest.mark.parametrize("attention_type", ["temporal", "spatial"])
def test_config2_layer_fwd_bwd_vs_oracle_full_size(attention_type):
    """cgat_layer_fwd / cgat_layer_bwd (the autograd path of the same layer) at the benched size."""
    ours, ref = _models(attention_type)
    layer = ours.net.hidden_layer
    x, _ = _batch()
    g = torch.Generator().manual_seed(5)
    dout = ((torch.rand(N, H, W, T, V, generator=g) - 0.5) / x.numel()).bfloat16()  # the size of a mean-loss gradient
    xf, df = x.float(), dout.float()
    keep = {}

    def run(m):
        keep["out"] = m(xf)
        return (keep["out"] * df).sum()

    _, total = _grads(ref, run, False)
    out_r = keep["out"].detach()
    _, emu = _grads(ref, run, True)
    abs_sum = {k: float((emu[k] - total[k]).pow(2).mean().sqrt()) for k in total}
    for p in layer.parameters():
        p.grad = None
    out_o = layer(x.to(DEV))
    out_o.backward(dout.to(DEV))
    torch.cuda.synchronize()
    # output: rtol 2e-2 + one bf16 ulp at the tensor's scale (the stored value is rounded to bf16)
    err = (out_o.float().cpu() - out_r).abs()
    tol = 2e-2 * out_r.abs() + BF16_ULP * float(out_r.abs().max())
    assert (err <= tol).all(), f"out: worst err {float(err.max()):.3e}, {int((err > tol).sum())} elements outside"
    for k, p in layer.named_parameters():
        _assert_grad(k, p.grad, total[k], abs_sum[k])


def test_config2_train_kernel_large_magnitude_inputs():
    """Range check of the packed-fp16 attention math (fp16 overflows at 65504, bf16 does not): inputs x 100 drive the
    projected features to O(100) and the attention logits past 50.  The paired-half kernel must notice (its range guard:
    score halves beyond 8 / non-finite sums, csrc/layer_fused.cu LF_SMAX), the fp32 instantiation re-runs the step inside
    the same launch sequence, and the step must match the oracle at the usual bar."""
    from cgat.train_step import TrainStep

    ours, ref = _models("temporal", seed=3)
    with torch.no_grad():  # make the logits large as well: scale the attention vectors
        for m in ours.net.hidden_layer.stream.attentions:
            m.a.mul_(4.0)
    ref.load_state_dict(ours.net.hidden_layer.state_dict())
    x, y = _batch(scale=100.0, seed=13)
    with torch.no_grad():
        out = ref(x[:1].float())
    assert float(out.abs().max()) > 20.0, "the inputs must drive the activations well past the O(1) range"
    loss_r, grads_r, abs_r, _ = _oracle_step(ref, x, y)
    ts = TrainStep(ours, x.to(DEV), y.to(DEV), lr=LR, use_graph=False)
    assert ts.fused_stream is not None
    ts._fwd_bwd()
    torch.cuda.synchronize()
    assert ts.range_guard_fired, "inputs x 100 must trip the range guard of the packed-fp16 kernel"
    assert torch.isfinite(ts.loss).all() and torch.isfinite(ts.flat_grad).all(), "overflow in the packed-fp16 math"
    assert abs(float(ts.loss[0]) - loss_r) <= 2e-2 * abs(loss_r), (float(ts.loss[0]), loss_r)
    for k, p in ours.net.hidden_layer.named_parameters():
        _assert_grad(k, p.grad, grads_r[k], abs_r[k])


def test_config2_moderately_large_inputs_stay_in_fp16_range_or_fall_back():
    """Inputs x 4 (scores of a few units): whichever kernel the guard picks, the step matches the oracle."""
    from cgat.train_step import TrainStep

    ours, ref = _models("temporal", seed=4)
    x, y = _batch(scale=4.0, seed=17)
    loss_r, grads_r, abs_r, _ = _oracle_step(ref, x, y)
    ts = TrainStep(ours, x.to(DEV), y.to(DEV), lr=LR, use_graph=False)
    ts._fwd_bwd()
    torch.cuda.synchronize()
    assert abs(float(ts.loss[0]) - loss_r) <= 2e-2 * abs(loss_r), (float(ts.loss[0]), loss_r, ts.range_guard_fired)
    for k, p in ours.net.hidden_layer.named_parameters():
        _assert_grad(k, p.grad, grads_r[k], abs_r[k])
