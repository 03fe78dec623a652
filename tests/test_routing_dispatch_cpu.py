"""Per-op unit tests against dense oracles (the reference has none of these - SURVEY.md section 4)."""
import pytest
import torch

from tutel_b200 import moe
from tutel_b200.ops import routing, dispatch


def _dense_locations(idx_ks, E):
    """Oracle: one-hot cumsum exactly as tutel/impls/fast_dispatch.py:155-171."""
    k, S = idx_ks.shape
    locs, acc = [], torch.zeros(E, dtype=torch.int64)
    for j in range(k):
        mask = torch.nn.functional.one_hot(idx_ks[j].long(), E)
        pos = torch.cumsum(mask, 0) - 1 + acc
        locs.append((pos * mask).sum(1))
        acc = acc + mask.sum(0)
    return torch.stack(locs).to(torch.int32), acc.to(torch.int32)


@pytest.mark.parametrize('S,E,k', [(10, 4, 2), (257, 7, 3), (64, 1, 1), (5, 8, 8)])
def test_locations_match_cumsum_oracle(S, E, k):
    torch.manual_seed(S + E)
    scores = torch.softmax(torch.randn(S, E), dim=1)
    crit, _ = routing.extract_critical(scores, top_k=k)
    want_loc, want_cnt = _dense_locations(crit.idx_ks, E)
    assert torch.equal(crit.loc_ks, want_loc)
    assert torch.equal(crit[5].to(torch.int32), want_cnt)


def test_capacity_rules_worked_example():
    # S=10, E=4, k=2
    torch.manual_seed(0)
    scores = torch.softmax(torch.randn(10, 4), dim=1)
    cap = lambda **kw: routing.extract_critical(scores, top_k=2, **kw)[0][4]
    max_count = int(routing.extract_critical(scores, top_k=2)[0][5].max())
    assert cap(capacity_factor=1.0) == 6
    assert cap(capacity_factor=2.0) == 12
    assert cap(capacity_factor=0.0) == max_count
    assert cap(capacity_factor=-0.5) == min(max_count, 2)
    assert cap(capacity_factor=-4.0) == min(max_count, 24)
    assert cap(capacity_factor=1.0, alignment=4) == 8


def test_gate_normalisation_and_l_aux():
    torch.manual_seed(1)
    scores = torch.softmax(torch.randn(33, 5), dim=1)
    crit, l_aux = routing.extract_critical(scores, top_k=2)
    g = torch.stack(crit[3])
    assert torch.allclose(g.sum(0), torch.ones(33), atol=1e-6)
    # gshard loss oracle
    top1 = scores.argmax(1)
    ce = torch.bincount(top1, minlength=5).float() * 5 / 33
    want = (scores.sum(0) * ce).sum() / 33
    assert torch.allclose(l_aux, want, rtol=1e-5)
    crit2, _ = routing.extract_critical(scores, top_k=2, normalize_gate=False)
    assert torch.allclose(torch.stack(crit2[3]), scores.gather(1, torch.topk(scores, 2, 1).indices).t())


def test_batch_prioritized_routing_orders_by_confidence():
    torch.manual_seed(2)
    scores = torch.softmax(torch.randn(40, 3) * 3, dim=1)
    crit, _ = routing.extract_critical(scores, top_k=1, batch_prioritized_routing=True, capacity_factor=0.5)
    loc, idx, C = crit.loc_ks[0], crit.idx_ks[0], crit[4]
    conf = scores.max(1)[0]
    for e in range(3):
        members = torch.nonzero(idx == e).view(-1)
        order = members[torch.argsort(loc[members])]
        assert torch.all(conf[order][:-1] >= conf[order][1:] - 1e-7)   # most confident tokens get the first slots
    assert int(crit[5].sum()) == 40                                   # true counts (the reference's are wrong here)


def _dense_moe_oracle(x, crit, is_postscore, fn):
    """einsum-style oracle: out[s] = sum_j w * fn(buf)[slot]."""
    E, C = crit[0], crit[4]
    S, M = x.shape
    buf = torch.zeros(E, C, M, dtype=x.dtype)
    for j in range(len(crit[1])):
        for s in range(S):
            e, l = int(crit[1][j][s]), int(crit[2][j][s])
            if l < C:
                buf[e, l] = x[s] * (1.0 if is_postscore else crit[3][j][s])
    y = fn(buf)
    out = torch.zeros(S, y.shape[-1], dtype=x.dtype)
    for j in range(len(crit[1])):
        for s in range(S):
            e, l = int(crit[1][j][s]), int(crit[2][j][s])
            if l < C:
                out[s] += y[e, l] * (crit[3][j][s] if is_postscore else 1.0)
    return buf, out


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('is_postscore', [True, False])
def test_encode_decode_forward_backward_vs_oracle(dtype, is_postscore):
    torch.manual_seed(3)
    S, E, M, k = 37, 5, 16, 2
    logits = torch.randn(S, E, dtype=dtype, requires_grad=True)
    x = torch.randn(S, M, dtype=dtype, requires_grad=True)
    w = torch.randn(E, M, M, dtype=dtype)

    def run(use_lib):
        scores = torch.softmax(logits, dim=1)
        crit, _ = routing.extract_critical(scores, top_k=k, capacity_factor=0.7)
        fn = lambda b: torch.tanh(torch.einsum('ecm,emn->ecn', b, w))
        if use_lib:
            buf = moe.fast_encode(x, crit, is_postscore)
            out = moe.fast_decode(fn(buf), crit, is_postscore)
        else:
            buf, out = _dense_moe_oracle(x, crit, is_postscore, fn)
        loss = (out * torch.arange(out.numel(), dtype=dtype).view_as(out)).sum()
        gx, gl = torch.autograd.grad(loss, [x, logits])
        return buf.detach(), out.detach(), gx, gl

    a, b = run(True), run(False)
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=1e-9, atol=1e-9)
    for u, v in zip(a, b):
        assert torch.allclose(u, v, **tol)


def test_half_precision_cpu_fallback_path():
    torch.manual_seed(4)
    x = torch.randn(20, 8).to(torch.bfloat16)
    scores = torch.softmax(torch.randn(20, 4), dim=1)
    crit, _ = routing.extract_critical(scores, top_k=2)
    buf = moe.fast_encode(x, crit)
    assert buf.dtype == torch.bfloat16 and buf.shape == (4, crit[4], 8)
    out = moe.fast_decode(buf, crit)
    kept = torch.stack([(crit[2][j] < crit[4]).float() * crit[3][j] for j in range(2)]).sum(0)
    assert torch.allclose(out.float(), x.float() * kept.unsqueeze(1), atol=3e-2)


def test_plain_tuple_critical_data_is_accepted():
    torch.manual_seed(5)
    x = torch.randn(12, 6)
    scores = torch.softmax(torch.randn(12, 3), dim=1)
    crit, _ = routing.extract_critical(scores, top_k=2)
    plain = tuple(crit)
    assert torch.equal(moe.fast_encode(x, plain), moe.fast_encode(x, crit))


def test_fast_cumsum_sub_one():
    m = (torch.rand(50, 6) > 0.5).int()
    assert torch.equal(moe.fast_cumsum_sub_one(m), torch.cumsum(m, 0) - 1)
    with pytest.raises(Exception):
        moe.fast_cumsum_sub_one(m, dim=1)


@pytest.mark.parametrize('k,normalize', [(1, True), (2, True), (2, False), (4, True)])
def test_fused_topk_gate_formulas_match_autograd(k, normalize):
    """ops/gating.FusedTopKGate (closed-form backward, same math as the CUDA kernels) vs autograd of the unfused ops."""
    from tutel_b200.models import losses
    from tutel_b200.ops.gating import fused_topk_gate
    torch.manual_seed(0)
    S, E = 37, 10
    logits = torch.randn(S, E, requires_grad=True)
    wg = torch.randn(k, S)
    # reference: separate softmax / topk / normalisation / loss, differentiated by autograd
    p = torch.softmax(logits, dim=1)
    ti = torch.topk(p, k, dim=1).indices
    g = p.gather(1, ti).t()
    if normalize and k > 1:
        g = g / torch.clamp(g.sum(dim=0, keepdim=True), min=torch.finfo(g.dtype).eps)
    l_ref = losses.gshard_loss(p, ti)
    ((g * wg).sum() + 3.0 * l_ref).backward()
    ref_grad = logits.grad.clone()
    logits.grad = None
    idx, gates, l_aux, top1 = fused_topk_gate(logits, k, normalize, True)
    assert torch.equal(idx.long(), ti.t()) and torch.allclose(gates, g.detach(), atol=1e-6)
    assert torch.allclose(l_aux, l_ref.detach(), atol=1e-6) and torch.allclose(top1, p.max(dim=1)[0].detach(), atol=1e-6)
    ((gates * wg).sum() + 3.0 * l_aux).backward()
    assert torch.allclose(logits.grad, ref_grad, atol=2e-6, rtol=1e-4)


def test_layer_with_fused_gate_matches_default_path(monkeypatch):
    from tutel_b200 import moe

    def run(fused, bpr):
        monkeypatch.setenv('TUTEL_B200_FUSED_GATE', '1' if fused else '0')
        torch.manual_seed(3)
        layer = moe.moe_layer(gate_type={'type': 'top', 'k': 2, 'capacity_factor': 1.5}, model_dim=16,
                              experts={'type': 'ffn', 'num_experts_per_device': 4, 'hidden_size_per_expert': 32},
                              batch_prioritized_routing=bpr)
        x = torch.randn(48, 16, requires_grad=True)
        y = layer(x)
        (y.pow(2).mean() + 0.1 * y.l_aux).backward()
        return y.detach(), y.l_aux.detach(), x.grad.clone(), layer.gates[0].wg.weight.grad.clone()

    for bpr in (False, True):
        a, b = run(True, bpr), run(False, bpr)
        for u, v in zip(a, b):
            assert torch.allclose(u, v, atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize('E,Sh,r', [(2, 4, 1), (2, 4, 2), (2, 4, 4), (1, 2, 2), (3, 2, 1)])
def test_virtual_geometry_of_sharded_experts_matches_repeat_view_sum(E, Sh, r):
    """parallel/fused.py expresses experts sharded over Sh GPUs with r token copies as one virtual expert per GPU.  Emulated
    with plain indexing, it must move exactly the rows the generic path moves with repeat / view / all-to-all / sum
    (models/moe_layer.py; reference tutel/impls/moe_layer.py:331-357)."""
    from tutel_b200.ops.dispatch import DispatchPlan, raw_decode, raw_encode
    from tutel_b200.ops.routing import _locations
    from tutel_b200.parallel.fused import _virtual_plan
    torch.manual_seed(E * 100 + Sh * 10 + r)
    S, k, M = 40, min(2, E), 6
    W = E * Sh
    scores = torch.rand(S, E)
    idx = torch.topk(scores, k, dim=1).indices.t().contiguous().to(torch.int32)
    loc, counts = _locations(idx, E)
    C = 16                                    # some tokens are dropped; C * r divisible by Sh
    plan = DispatchPlan(E, C, idx, loc.to(torch.int32))
    x = torch.randn(S, M)
    gates = torch.rand(k, S)
    gpu_scale = torch.arange(1, W + 1, dtype=torch.float32).view(W, 1, 1)      # "expert compute" that differs per GPU

    # generic path
    enc = raw_encode(x, None, plan).view(E, C, M)
    sent = enc.repeat(1, r, 1).view(W, -1, M)
    back = (sent * gpu_scale).view(E, r, -1, M).sum(dim=1)
    want = raw_decode(back.reshape(E * C, M), gates, plan)

    # virtual geometry
    vp = _virtual_plan(plan, E, Sh, r)
    Cv = C * r // Sh
    slot = vp.slot_src.view(W, Cv).long()
    recv = torch.where((slot >= 0).unsqueeze(-1), x[(slot.clamp(min=0) // k)], torch.zeros(()))
    out_buf = (recv * gpu_scale).reshape(W * Cv, M)
    vplan = DispatchPlan(W, Cv, vp.idx_ks, vp.loc_ks)
    got = raw_decode(out_buf, gates.repeat(r, 1), vplan)
    assert torch.allclose(got, want, atol=1e-5)
