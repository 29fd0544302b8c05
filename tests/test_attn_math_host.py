"""The per-pixel attention arithmetic of the CUDA kernels (csrc/attn_math.cuh), compiled for the host,
against the oracle (oracle/spec.py attention_core + autograd).  CPU only."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import spec

F = ctypes.POINTER(ctypes.c_float)
U64 = ctypes.POINTER(ctypes.c_uint64)


def fp(t):
    return t.contiguous().numpy().ctypes.data_as(F) if t is not None else None


def maskrows(mask):
    rows = np.zeros(mask.shape[0], dtype=np.uint64)
    for i in range(mask.shape[0]):
        for j in range(mask.shape[1]):
            if mask[i, j]:
                rows[i] |= np.uint64(1) << np.uint64(j)
    return rows


CASES = [(6, 4, 4), (4, 6, 6), (8, 4, 4)]


def make(nodes, ci, co, P, seed, masked):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(2, P, nodes, ci, generator=g)
    W = (torch.rand(ci, co, generator=g) - 0.5) * 2
    a = (torch.rand(2 * co, generator=g) - 0.5) * 2
    B = torch.rand(nodes, nodes, generator=g) * 0.5
    mask = torch.ones(nodes, nodes, dtype=torch.uint8)
    if masked:
        mask = (torch.rand(nodes, nodes, generator=g) < 0.5).to(torch.uint8)
        mask |= torch.eye(nodes, dtype=torch.uint8)
        mask[-1] = 0  # a fully masked row: uniform attention (softmax of equal logits)
    dZ = torch.rand(2, P, nodes, co, generator=g) - 0.5
    return X, W, a, B, mask, dZ


@pytest.mark.parametrize("nodes,ci,co", CASES)
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("axis", ["neighbour", "pixel"])
def test_forward_backward_match_oracle(host_harness, nodes, ci, co, masked, axis):
    hh = host_harness
    P = 37
    X, W, a, B, mask, dZ = make(nodes, ci, co, P, 369 + nodes, masked)
    Xr, Wr, ar, Br = (t.clone().requires_grad_() for t in (X, W, a, B))
    adj = spec.adjacency_norm(Br)
    adj.retain_grad()
    z = spec.attention_core(Xr @ Wr, ar, adj, alpha=0.2, softmax_axis=axis, mask=mask, apply_elu=False)
    z.backward(dZ)

    rows = maskrows(mask.numpy())
    adj_c = adj.detach().contiguous()
    pix = axis == "pixel"
    zs, dXs = [], []
    gW = np.zeros((ci, co), np.float32)
    ga = np.zeros(2 * co, np.float32)
    gadj = np.zeros((nodes, nodes), np.float32)
    for n in range(2):  # statistics are per sample
        Xn = X[n].contiguous()
        st_max = st_rinv = st_dot = None
        if pix:
            e = torch.zeros(P, nodes, nodes)
            assert hh.hh_logits(nodes, ci, co, P, fp(Xn), fp(W), fp(a), rows.ctypes.data_as(U64),
                                ctypes.c_float(0.2), fp(e)) == 0
            st_max = e.max(dim=0).values.contiguous()
            st_rinv = (1.0 / torch.exp(e - st_max).sum(0)).contiguous()
            st_dot = torch.zeros(nodes, nodes)
            assert hh.hh_bwd(nodes, ci, co, 1, 1, P, fp(Xn), fp(dZ[n]), fp(W), fp(a), fp(adj_c),
                             rows.ctypes.data_as(U64), ctypes.c_float(0.2), fp(st_max), fp(st_rinv), None, None, None,
                             None, None, fp(st_dot)) == 0
        zo = torch.zeros(P, nodes, co)
        assert hh.hh_fwd(nodes, ci, co, int(pix), P, fp(Xn), fp(W), fp(a), fp(adj_c), rows.ctypes.data_as(U64),
                         ctypes.c_float(0.2), fp(st_max), fp(st_rinv), fp(zo)) == 0
        zs.append(zo)
        dXo = torch.zeros(P, nodes, ci)
        assert hh.hh_bwd(nodes, ci, co, int(pix), 0, P, fp(Xn), fp(dZ[n]), fp(W), fp(a), fp(adj_c),
                         rows.ctypes.data_as(U64), ctypes.c_float(0.2), fp(st_max), fp(st_rinv), fp(st_dot), fp(dXo),
                         gW.ctypes.data_as(F), ga.ctypes.data_as(F), gadj.ctypes.data_as(F), None) == 0
        dXs.append(dXo)
    torch.testing.assert_close(torch.stack(zs), z.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(torch.stack(dXs), Xr.grad, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(torch.from_numpy(gW), Wr.grad, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(torch.from_numpy(ga), ar.grad, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(torch.from_numpy(gadj), adj.grad, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("nodes,ci,co", CASES)
@pytest.mark.parametrize("masked", [False, True])
def test_restructured_neighbour_math_matches_oracle(host_harness, nodes, ci, co, masked):
    """attn_nb_forward / attn_nb_backward (M = adj^T.att form used by the fused layer kernels) vs autograd."""
    hh = host_harness
    P = 41
    X, W, a, B, mask, dZ = make(nodes, ci, co, P, 1234 + nodes, masked)
    Xr, Wr, ar, Br = (t.clone().requires_grad_() for t in (X, W, a, B))
    adj = spec.adjacency_norm(Br)
    adj.retain_grad()
    z = spec.attention_core(Xr @ Wr, ar, adj, alpha=0.2, softmax_axis="neighbour", mask=mask, apply_elu=False)
    z.backward(dZ)
    rows = maskrows(mask.numpy())
    adj_c = adj.detach().contiguous()
    n_pix = 2 * P
    zo = torch.zeros(n_pix, nodes, co)
    dXo = torch.zeros(n_pix, nodes, ci)
    gW = np.zeros((ci, co), np.float32)
    ga = np.zeros(2 * co, np.float32)
    gadj = np.zeros((nodes, nodes), np.float32)
    # the unmasked instantiation is only valid for an all-ones mask
    assert hh.hh_nb(nodes, ci, co, int(masked), n_pix, fp(X.reshape(n_pix, nodes, ci)), fp(dZ.reshape(n_pix, nodes, co)),
                    fp(W), fp(a), fp(adj_c), rows.ctypes.data_as(U64), ctypes.c_float(0.2), fp(zo), fp(dXo),
                    gW.ctypes.data_as(F), ga.ctypes.data_as(F), gadj.ctypes.data_as(F)) == 0
    torch.testing.assert_close(zo.reshape(2, P, nodes, co), z.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dXo.reshape(2, P, nodes, ci), Xr.grad, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(torch.from_numpy(gW), Wr.grad, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(torch.from_numpy(ga), ar.grad, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(torch.from_numpy(gadj), adj.grad, rtol=1e-4, atol=2e-4)
