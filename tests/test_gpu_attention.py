"""GPU parity tests of the fused attention path (through the C ABI) against the golden vectors of the live
reference and against the oracle (oracle/spec.py).  fp32: rtol 1e-4; bf16: rtol 2e-2 (north-star tolerances)."""
import pytest
import torch

from oracle import spec
from util import close, close_frac, golden, sd_of

pytestmark = pytest.mark.gpu

DEV = "cuda"


# ---------------------------------------------------------------------------------------------------
def test_adjacency_norm_fwd_bwd_vs_reference_golden():
    import ctypes

    from cgat import _lib
    from cgat.functional import adjacency_norm

    fx = golden("adjacency")
    for V in (4, 6, 32):
        B = fx[f"V{V}.B"].to(DEV)[None]
        for transpose in (False, True):
            ah = adjacency_norm(B, transpose)[0]
            ref = fx[f"V{V}.A_hat"]
            close(ah, ref.t() if transpose else ref, msg=f"A_hat V={V} T={transpose}")
            g = fx[f"V{V}.g"]
            gk = (g.t() if transpose else g).contiguous().to(DEV)[None]
            gB = torch.empty_like(B)
            _lib.check(_lib.lib().cgat_adj_norm_bwd(_lib.ptr(B), _lib.ptr(gk), _lib.ptr(gB), 1, V, int(transpose),
                                                    _lib.stream()), "adj bwd")
            close(gB[0], fx[f"V{V}.grad_B"], atol=1e-5, msg=f"grad_B V={V} T={transpose}")


@pytest.mark.parametrize("name", ["gat2d_layer", "gat2d_layer_init"])
def test_gat2d_layer_vs_reference_golden(name):
    from cgat.layers import GraphAttentionLayer2D

    fx = golden(name)
    N, P, T, V = fx["h"].shape
    lay = GraphAttentionLayer2D(T, T, V, fx["alpha"]).to(DEV)
    lay.load_state_dict(sd_of(fx))
    h = fx["h"].to(DEV).requires_grad_()
    out = lay(h)
    close(out, fx["out"], msg="out")
    out.backward(fx["g"].to(DEV))
    close(h.grad, fx["grad_in.h"], atol=1e-5, msg="dh")
    close(lay.W.grad, fx["grad.W"], atol=1e-4, msg="dW")
    close(lay.a.grad, fx["grad.a"], atol=1e-4, msg="da")
    close(lay.B.grad, fx["grad.B"], atol=1e-4, msg="dB")


def test_baseline2d_model_vs_reference_golden():
    from convolutional_gat.baseline_model import BaselineModel2D

    fx = golden("baseline2d_model")
    N, H, W, T, V = fx["x"].shape
    model = BaselineModel2D(image_width=W, image_height=H, n_vertices=V).to(DEV)
    model.load_state_dict(sd_of(fx))  # same keys as the reference's state_dict
    x = fx["x"].to(DEV).requires_grad_()
    out = model(x)
    close(out, fx["out"], msg="out")
    out.backward(fx["g"].to(DEV))
    close(x.grad, fx["grad_in.x"], atol=1e-5, msg="dx")
    for k, p in model.named_parameters():
        close(p.grad, fx[f"grad.{k}"], atol=1e-4, msg=f"d{k}")


def test_gat1d_layer_vs_reference_golden():
    from cgat.layers import GraphAttentionLayer

    fx = golden("gat1d_layer")
    N, V, F_ = fx["h"].shape
    lay = GraphAttentionLayer(F_, F_, V, fx["alpha"]).to(DEV)
    lay.load_state_dict(sd_of(fx))
    h = fx["h"].to(DEV).requires_grad_()
    out = lay(h)
    close(out, fx["out"], msg="out")
    out.backward(fx["g"].to(DEV))
    close(h.grad, fx["grad_in.h"], atol=1e-5, msg="dh")
    close(lay.W.grad, fx["grad.W"], atol=1e-4, msg="dW")
    close(lay.a.grad, fx["grad.a"], atol=1e-4, msg="da")
    close(lay.B.grad, fx["grad.B"], atol=1e-4, msg="dB")


def test_baseline1d_model_vs_reference_golden():
    from convolutional_gat.baseline_model import BaselineModel

    fx = golden("baseline1d_model")
    N, H, W, T, V = fx["x"].shape
    model = BaselineModel(image_width=W, image_height=H, n_vertices=V).to(DEV)
    model.load_state_dict(sd_of(fx))
    x = fx["x"].to(DEV).requires_grad_()
    out = model(x)
    close(out, fx["out"], msg="out")
    out.backward(fx["g"].to(DEV))
    close(x.grad, fx["grad_in.x"], atol=1e-5, msg="dx")
    for k, p in model.named_parameters():
        close(p.grad, fx[f"grad.{k}"], atol=1e-4, msg=f"d{k}")


# ---------------------------------------------------------------------------------------------------
def _pair(type_, mapping, heads, merge, axis, masked, seed, T=4, V=6):
    """(ours on GPU, spec oracle on CPU) with identical parameters."""
    from cgat.layers import GATMultiHead3D

    torch.manual_seed(seed)
    ref = spec.SpecGATMultiHead3D(T, T, 0.2, heads, type_=type_, mapping_type=mapping, n_vertices=V,
                                  softmax_axis=axis, head_merge=merge)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.endswith(".B"):
                p.add_(torch.rand_like(p) * 0.3)
    ours = GATMultiHead3D(T, T, 0.2, heads, type_=type_, mapping_type=mapping, n_vertices=V, softmax_axis=axis,
                          head_merge=merge)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV)
    if masked:
        for (so, sr) in zip([m for m in ours.modules() if hasattr(m, "adj_mask")],
                            [m for m in ref.modules() if hasattr(m, "adj_mask")]):
            n = sr.adj_mask.shape[0]
            m = (torch.rand(n, n) < 0.5).to(torch.uint8) | torch.eye(n, dtype=torch.uint8)
            m[-1] = 0
            sr.adj_mask.copy_(m)
            so.adj_mask.copy_(m.to(DEV))
            assert torch.equal(so.adj_mask.cpu(), sr.adj_mask)  # bit-exact mask
    return ours, ref


def _check(ours, ref, x, dtype, rtol, atol, gatol):
    xr = x.clone().requires_grad_()
    out_r = ref(xr)
    g = torch.rand_like(out_r) - 0.5
    out_r.backward(g)
    xo = x.to(DEV, dtype).requires_grad_()
    out_o = ours(xo)
    assert out_o.dtype == dtype and out_o.shape == out_r.shape
    # absolute slack is relative to the tensor's scale (bf16 itself rounds a value of magnitude m by m * 2^-9)
    close(out_o, out_r.detach(), rtol=rtol, atol=atol * max(1.0, out_r.abs().max().item()), msg="out")
    out_o.backward(g.to(DEV, dtype))
    if dtype == torch.bfloat16:  # see util.close_frac: kink flips of LeakyReLU' on isolated pixels
        close_frac(xo.grad, xr.grad, rtol=rtol, atol=atol * max(1.0, xr.grad.abs().max().item()), msg="dx")
    else:
        close(xo.grad, xr.grad, rtol=rtol, atol=atol * max(1.0, xr.grad.abs().max().item()), msg="dx")
    pr = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        assert p.grad is not None, k
        close(p.grad, pr[k].grad, rtol=rtol, atol=gatol * max(1.0, pr[k].grad.abs().max().item()), msg=f"d{k}")


@pytest.mark.parametrize("type_", ["spatial", "temporal", "multi_stream"])
@pytest.mark.parametrize("heads,merge", [(1, "mean"), (3, "mean"), (3, "concat")])
@pytest.mark.parametrize("masked", [False, True])
def test_gat3d_linear_neighbour_fp32(type_, heads, merge, masked):
    if type_ == "multi_stream" and merge == "concat":
        pytest.skip("streams concatenate on different axes")
    ours, ref = _pair(type_, "linear", heads, merge, "neighbour", masked, seed=11)
    x = torch.rand(2, 13, 11, 4, 6)  # 143 pixels/sample: tiles straddle samples, ragged last tile
    _check(ours, ref, x, torch.float32, 1e-4, 1e-5, 1e-4)


@pytest.mark.parametrize("type_", ["spatial", "temporal"])
def test_gat3d_linear_pixel_softmax_fp32(type_):
    ours, ref = _pair(type_, "linear", 2, "concat", "pixel", True, seed=12)
    x = torch.rand(3, 10, 9, 4, 6)
    _check(ours, ref, x, torch.float32, 1e-4, 1e-5, 1e-4)


@pytest.mark.parametrize("type_", ["spatial", "temporal"])
@pytest.mark.parametrize("merge", ["mean", "concat"])
def test_gat3d_linear_bf16(type_, merge):
    ours, ref = _pair(type_, "linear", 3, merge, "neighbour", False, seed=13)
    x = torch.rand(4, 32, 32, 4, 6).bfloat16().float()  # bf16-representable inputs; oracle runs fp32 on them
    # parameter gradients are sums over pixels: LeakyReLU' kink flips on isolated pixels (util.close_frac) move them
    # by O(1/n_pix) each, hence the slightly wider absolute slack (relative to the largest gradient entry)
    _check(ours, ref, x, torch.bfloat16, 2e-2, 2e-2, 4e-2)


@pytest.mark.parametrize("type_", ["spatial", "temporal", "multi_stream"])
@pytest.mark.parametrize("heads,merge", [(1, "mean"), (3, "mean"), (3, "concat")])
def test_gat3d_conv_mapping_fp32(type_, heads, merge):
    if type_ == "multi_stream" and merge == "concat":
        pytest.skip("streams concatenate on different axes")
    ours, ref = _pair(type_, "conv", heads, merge, "neighbour", False, seed=14)
    x = torch.rand(2, 12, 10, 4, 6)
    _check(ours, ref, x, torch.float32, 1e-4, 2e-5, 1e-4)


def test_gat3d_conv_mapping_bf16():
    ours, ref = _pair("temporal", "conv", 3, "mean", "neighbour", False, seed=15)
    x = torch.rand(2, 16, 16, 4, 6).bfloat16().float()
    _check(ours, ref, x, torch.bfloat16, 2e-2, 2e-2, 3e-2)


def test_empty_batch_rejected():
    from cgat.layers import GATMultiHead3D

    layer = GATMultiHead3D(4, 4, 0.2, 1, type_="spatial", n_vertices=6).to(DEV)
    with pytest.raises(RuntimeError):
        layer(torch.rand(0, 4, 4, 4, 6, device=DEV))


def test_unsupported_shape_fails_loudly():
    from cgat.layers import GATMultiHead3D

    layer = GATMultiHead3D(4, 4, 0.2, 1, type_="spatial", n_vertices=65).to(DEV)  # generic kernel serves <= 64 nodes
    with pytest.raises(RuntimeError):
        layer(torch.rand(1, 4, 4, 4, 65, device=DEV))
    layer = GATMultiHead3D(4, 4, 0.2, 1, type_="spatial", n_vertices=5, softmax_axis="pixel").to(DEV)
    with pytest.raises(RuntimeError, match="not instantiated"):
        layer(torch.rand(1, 4, 4, 4, 5, device=DEV))


# ---------------------------------------------------------------------------------------------------
# many nodes (BASELINE config 4: V in {32, 64}): the row-of-threads-per-pixel kernels of attn_generic.cu
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,type_,mapping,heads,merge,masked", [
    (32, "spatial", "linear", 3, "mean", False),
    (32, "spatial", "linear", 2, "concat", True),
    (64, "spatial", "linear", 1, "mean", True),
    (10, "spatial", "linear", 2, "mean", True),     # a node count that is not a power of two
    (32, "spatial", "conv", 2, "mean", False),
    (5, "temporal", "linear", 2, "concat", False),  # nodes = 4 (T), channels = 5
])
def test_gat3d_many_nodes_fp32(V, type_, mapping, heads, merge, masked):
    ours, ref = _pair(type_, mapping, heads, merge, "neighbour", masked, seed=41, V=V)
    x = torch.rand(2, 9, 7, 4, V)
    _check(ours, ref, x, torch.float32, 1e-4, 2e-5, 2e-4)


def test_gat3d_many_nodes_bf16_stress_shape():
    """BASELINE config 4's attention shape at reduced batch: 128 x 128 pixels, V = 32 nodes, bf16."""
    ours, ref = _pair("spatial", "linear", 3, "mean", "neighbour", True, seed=42, V=32)
    x = torch.rand(1, 128, 128, 4, 32).bfloat16().float()
    _check(ours, ref, x, torch.bfloat16, 2e-2, 2e-2, 4e-2)


# ---------------------------------------------------------------------------------------------------
# full-size (BASELINE config 2: N=64, 64x64, T=4, V=6, bf16) properties that need no CPU oracle run
# ---------------------------------------------------------------------------------------------------
def test_full_size_tile_invariance_and_closed_form():
    from cgat.layers import GATMultiHead3D

    torch.manual_seed(369)
    layer = GATMultiHead3D(4, 4, 0.2, 3, type_="temporal", mapping_type="linear", n_vertices=6).to(DEV)
    x = torch.rand(64, 64, 64, 4, 6, device=DEV).bfloat16()
    with torch.no_grad():
        full = layer(x)
        # (1) samples are independent: any slice of the batch gives the same numbers, bit for bit
        part = layer(x[17:19].contiguous())
        assert torch.equal(full[17:19], part)
        # (2) pixels are independent: a spatial crop gives the same numbers for the cropped pixels
        crop = layer(x[:2, 5:37, 8:24].contiguous())
        assert torch.equal(full[:2, 5:37, 8:24], crop)
        # (3) closed form: with a = 0 the attention is uniform and with B at its init A_hat = I, so
        #     out[node] = ELU(mean_j Wh[j]) for every node -- checked against plain torch on the GPU
        for m in layer.stream.attentions:
            m.a.zero_()
        out = layer(x).float()
        xf = x.float()  # temporal: nodes = T (dim 3), channels = V (dim 4)
        want = 0
        for m in layer.stream.attentions:
            wh = xf @ m.W  # [N,H,W,T,V']
            want = want + torch.nn.functional.elu(wh.mean(dim=3, keepdim=True).expand_as(wh))
        want = want / 3
    close(out, want, rtol=2e-2, atol=2e-2, msg="closed form at full size")


def test_gat3d_smaat_unet_mapping_fp32():
    """``mapping_type="smaat_unet"`` (the third value the reference's call sites pass, model.py:21-42): the shared
    SmaAt-UNet per node feeds the attention kernels.  Eval mode (BatchNorm running statistics) against the oracle."""
    ours, ref = _pair("spatial", "smaat_unet", 1, "mean", "neighbour", False, seed=51)
    ours.eval()
    ref.eval()
    torch.manual_seed(4)
    x = torch.rand(1, 16, 16, 4, 6)
    xr = x.clone().requires_grad_()
    out_r = ref(xr)
    g = torch.rand_like(out_r) - 0.5
    out_r.backward(g)
    xo = x.to(DEV).requires_grad_()
    out_o = ours(xo)
    close(out_o, out_r.detach(), rtol=1e-3, atol=1e-4, msg="out")
    out_o.backward(g.to(DEV))
    close(xo.grad, xr.grad, rtol=1e-3, atol=1e-4 * max(1.0, xr.grad.abs().max().item()), msg="dx")
    pr = dict(ref.named_parameters())
    for k, p in ours.named_parameters():
        r = pr[k].grad
        close(p.grad, r, rtol=2e-3, atol=2e-4 * max(1.0, r.abs().max().item()), msg=f"d{k}")
