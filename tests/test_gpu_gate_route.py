"""Fused gate + routing kernels, column sums and the public column scan against plain PyTorch fp32 references
(csrc/gate_route.cu; run with `pytest -m gpu` on a B200)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def C():
    from tutel_b200.ops import backend
    return backend.require_ext()


def _torch_gate(logits, k, normalize):
    """Op-by-op reference: softmax, top-k, normalised gates, GShard loss (tutel/impls/fast_dispatch.py:143-176)."""
    S, E = logits.shape
    p = torch.softmax(logits.float(), dim=1)
    top, idx = torch.topk(p, k, dim=1)
    gates = top
    if normalize and k > 1:
        gates = top / torch.clamp(top.sum(dim=1, keepdim=True), min=float(torch.finfo(logits.dtype).eps))
    ce = torch.bincount(idx[:, 0], minlength=E).float()
    l_aux = (p.sum(0) * ce).sum() * (E / float(S * S))
    return p, idx, top, gates, l_aux


def _torch_locations(idx_ks, E):
    """Stable queue positions: all first choices in token order, then all second choices, ... """
    k, S = idx_ks.shape
    flat = idx_ks.reshape(-1).long()
    onehot = F.one_hot(flat, E)
    pos = torch.cumsum(onehot, dim=0) - 1
    return pos.gather(1, flat.unsqueeze(1)).view(k, S), onehot.sum(0)


@pytest.mark.parametrize('S,E,k,dtype', [(8192, 8, 2, torch.bfloat16), (777, 130, 4, torch.float32),
                                         (1000, 64, 1, torch.float16), (300, 512, 8, torch.float32)])
def test_gate_route_forward_matches_torch(C, S, E, k, dtype):
    torch.manual_seed(3)
    logits = (torch.randn(S, E, device='cuda') * 2).to(dtype)
    cap = k * ((S + E - 1) // E)
    scores, idx, top, gates, loc, counts, ce, l_aux, slot = C.gate_route_forward(logits, k, cap, True, float(torch.finfo(dtype).eps))
    p, ridx, rtop, rgates, rl = _torch_gate(logits, k, True)
    assert torch.allclose(scores, p, atol=2e-6)
    # ties between equal scores may be ordered differently by torch.topk: compare the selected VALUES, and ids where unique
    assert torch.allclose(top.t(), rtop, atol=2e-6)
    same = idx.t().long() == ridx
    assert same.float().mean() > 0.995
    assert torch.allclose(gates.t(), rgates, atol=1e-5, rtol=1e-4)
    assert torch.allclose(l_aux.float(), rl, rtol=2e-2 if dtype != torch.float32 else 1e-4)
    # routing of the kernel's own choices
    rloc, rcounts = _torch_locations(idx, E)
    assert torch.equal(loc.long(), rloc) and torch.equal(counts.long(), rcounts)
    assert torch.equal(ce, torch.bincount(idx[0].long(), minlength=E).float())
    # the inverse map: slot (e, l) -> token * k + choice, -1 where empty
    want = torch.full([E * cap], -1, dtype=torch.int32, device='cuda')
    valid = loc < cap
    tok = torch.arange(S, device='cuda', dtype=torch.int32).unsqueeze(0) * k + torch.arange(k, device='cuda', dtype=torch.int32).unsqueeze(1)
    want[(idx.long() * cap + loc.long())[valid]] = tok[valid]
    assert torch.equal(slot, want)


@pytest.mark.parametrize('E,k,normalize,dtype', [(8, 2, True, torch.bfloat16), (130, 4, True, torch.float32),
                                                (64, 3, False, torch.float32), (16, 1, True, torch.float16)])
def test_gate_route_autograd_matches_torch_autograd(E, k, normalize, dtype):
    """One-launch backward (normalisation + top-k selection + softmax + loss) vs autograd of the op-by-op formulation."""
    from tutel_b200.ops.gating import fused_gate_route
    torch.manual_seed(6)
    S = 1111
    base = torch.randn(S, E, device='cuda').to(dtype)
    wg = torch.randn(k, S, device='cuda')
    a = base.clone().requires_grad_(True)
    idx, loc, gates, l_aux, counts, top1, slot = fused_gate_route(a, k, normalize, 0)
    ((gates * wg).sum() + 2.0 * l_aux.float()).backward()
    b = base.float().clone().requires_grad_(True)
    p = torch.softmax(b, dim=1)
    rtop = p.gather(1, idx.t().long())                       # same selection as the kernel
    rg = rtop / torch.clamp(rtop.sum(1, keepdim=True), min=float(torch.finfo(dtype).eps)) if (normalize and k > 1) else rtop
    ce = torch.bincount(idx[0].long(), minlength=E).float()
    rl = (p.sum(0) * ce).sum() * (E / float(S * S))
    ((rg.t() * wg).sum() + 2.0 * rl).backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert slot is None and torch.allclose(gates, rg.t(), atol=1e-5)
    assert torch.allclose(a.grad.float(), b.grad, atol=tol * b.grad.abs().max().item() + 1e-7, rtol=tol)


def test_layer_fused_gate_matches_op_by_op(monkeypatch):
    """The whole layer with the fused CUDA gate/route vs TUTEL_B200_FUSED_GATE=0 (fp32: tight tolerance)."""
    from tutel_b200 import moe
    outs = []
    for mode in ('0', 'auto'):
        monkeypatch.setenv('TUTEL_B200_FUSED_GATE', mode)
        torch.manual_seed(0)
        layer = moe.moe_layer(gate_type={'type': 'top', 'k': 2, 'capacity_factor': 1.25}, model_dim=64,
                              experts={'type': 'ffn', 'num_experts_per_device': 4, 'hidden_size_per_expert': 128,
                                       'activation_fn': lambda t: F.relu(t)}, seeds=(1, 1, 1)).cuda()
        x = torch.randn(6, 50, 64, device='cuda', requires_grad=True)
        y = layer(x)
        (y.pow(2).mean() + 0.1 * y.l_aux).backward()
        outs.append((y.detach(), y.l_aux.detach(), x.grad.clone(), layer.gates[0].wg.weight.grad.clone(),
                     layer.dispatch_count.clone()))
    for u, v in zip(outs[0], outs[1]):
        assert torch.allclose(u.float(), v.float(), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize('G,T,N,dtype', [(8, 2048, 4096, torch.bfloat16), (1, 333, 264, torch.float32),
                                         (3, 1000, 520, torch.float16), (2, 16384, 128, torch.bfloat16)])
def test_grouped_colsum(C, G, T, N, dtype):
    torch.manual_seed(5)
    x = torch.randn(G, T, N, device='cuda').to(dtype)
    out = C.grouped_colsum(x)
    ref = x.float().sum(dim=1)
    assert out.dtype == dtype and torch.allclose(out.float(), ref, atol=2e-2 * T ** 0.5 if dtype != torch.float32 else 1e-3, rtol=1e-2)
    xs = x[:, : T // 2]                                      # strided groups (a view)
    assert torch.allclose(C.grouped_colsum(xs).float(), xs.float().sum(dim=1), atol=2e-2 * T ** 0.5, rtol=1e-2)


@pytest.mark.parametrize('S,E', [(8192, 8), (1000, 130), (33, 1), (70000, 3), (4097, 2048)])
def test_fast_cumsum_sub_one(S, E):
    from tutel_b200 import moe
    torch.manual_seed(8)
    mask = (torch.rand(S, E, device='cuda') < 0.3).to(torch.int64)
    out = moe.fast_cumsum_sub_one(mask)
    assert torch.equal(out.long(), torch.cumsum(mask, dim=0) - 1)
