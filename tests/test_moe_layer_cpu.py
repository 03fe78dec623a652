"""MoE layer behaviour on CPU: options, per-forward overrides, state dict compatibility, custom gate/expert."""
import copy

import pytest
import torch
import torch.nn.functional as F

from tutel_b200 import moe


def make(E=4, M=16, H=32, k=2, **kw):
    torch.manual_seed(0)
    experts = {'type': 'ffn', 'num_experts_per_device': E, 'hidden_size_per_expert': H, 'activation_fn': lambda x: F.relu(x)}
    experts.update(kw.pop('experts', {}))
    return moe.moe_layer(gate_type=kw.pop('gate_type', {'type': 'top', 'k': k}), model_dim=M, experts=experts,
                         seeds=(1, 2, 3), **kw)


def dense_reference(layer, x, k):
    """Every token through its top-k experts with no capacity limit."""
    S = x.reshape(-1, x.shape[-1])
    scores = torch.softmax(layer.gates[0](S), dim=1)
    tv, ti = torch.topk(scores, k, dim=1)
    tv = tv / tv.sum(1, keepdim=True)
    e = layer.experts
    out = torch.zeros(S.shape[0], e.output_dim)
    for j in range(k):
        for s in range(S.shape[0]):
            i = int(ti[s, j])
            h = F.relu(S[s] @ e.batched_fc1_w[i].t() + e.batched_fc1_bias[i])
            out[s] += tv[s, j] * (h @ e.batched_fc2_w[i] + e.batched_fc2_bias[i])
    return out.view(*x.shape[:-1], -1)


def test_forward_matches_dense_reference_when_nothing_is_dropped():
    layer = make(gate_type={'type': 'top', 'k': 2, 'capacity_factor': 0})   # dropless (a per-forward 0 means "default")
    x = torch.randn(3, 7, 16)
    y = layer(x)
    assert torch.allclose(y, dense_reference(layer, x, 2), atol=1e-5)
    assert y.l_aux is not None and layer.l_aux is y.l_aux
    assert int(layer.dispatch_count.sum()) == 3 * 7 * 2


def test_state_dict_keys_shapes_and_tolerant_loading():
    layer = make(E=2, M=8, H=12, experts={'output_dim': 6})
    sd = layer.state_dict()
    assert set(sd) == {'_num_global_experts', 'experts.batched_fc1_w', 'experts.batched_fc2_w', 'experts.batched_fc1_bias',
                       'experts.batched_fc2_bias', 'gates.0.wg.weight'}
    assert sd['experts.batched_fc1_w'].shape == (2, 12, 8) and sd['experts.batched_fc2_w'].shape == (2, 12, 6)
    assert sd['experts.batched_fc2_bias'].shape == (2, 6) and int(sd['_num_global_experts']) == 2
    other = make(E=2, M=8, H=12, experts={'output_dim': 6})
    legacy = {k: v.clone() for k, v in sd.items() if k != '_num_global_experts' and k != 'experts.batched_fc1_bias'}
    legacy['experts.batched_fc1_w'] = legacy['experts.batched_fc1_w'].reshape(2, -1)   # same numel, other shape
    other.load_state_dict(legacy)
    assert torch.equal(other.experts.batched_fc1_w, layer.experts.batched_fc1_w)
    assert torch.count_nonzero(other.experts.batched_fc1_bias) == 0                     # missing -> zero filled
    bad = dict(sd)
    bad['_num_global_experts'] = torch.tensor(5)
    with pytest.raises(AssertionError):
        make(E=2, M=8, H=12, experts={'output_dim': 6}).load_state_dict(bad)


def test_per_forward_overrides_and_multiple_gates():
    layer = make(gate_type=[{'type': 'top', 'k': 1}, {'type': 'top', 'k': 2, 'capacity_factor': 2.0}])
    x = torch.randn(2, 9, 16)
    y0 = layer(x, gate_index=0)
    y1 = layer(x, gate_index=1)
    assert y0.shape == y1.shape == x.shape and not torch.allclose(y0, y1)
    y_top1 = layer(x, gate_index=1, top_k=1, capacity_factor=4.0)
    assert y_top1.shape == x.shape
    with pytest.raises(Exception):
        moe.moe_layer(gate_type={'type': 'top', 'k': 1}, model_dim=16, experts={'type': 'ffn', 'hidden_size_per_expert': 8}, bogus=1)
    assert moe.moe_layer(gate_type='Top2Gate', model_dim=16, experts={'type': 'ffn', 'hidden_size_per_expert': 8},
                         pad_samples=True).gates[0].top_k == 1   # k is clamped to the single expert


def test_reserve_dims_and_result_func_and_skip_flags(monkeypatch):
    layer = make(M=24, result_func=lambda t: t * 2)
    x = torch.randn(5, 4, 6)          # model_dim 24 = 4*6 spread over 2 trailing dims
    y = layer(x, reserve_dims=2)
    assert y.shape == x.shape
    monkeypatch.setenv('SKIP_MOE', '1')
    ident = make(M=24)
    assert torch.equal(ident(x.view(5, 24)), x.view(5, 24))


def test_prescore_and_postscore_both_train():
    for post in (True, False):
        layer = make(is_postscore=post)
        x = torch.randn(4, 6, 16, requires_grad=True)
        layer(x).pow(2).sum().backward()
        assert x.grad.abs().sum() > 0 and layer.gates[0].wg.weight.grad.abs().sum() > 0


def test_cosine_gate_and_load_importance_loss_and_noise():
    layer = make(gate_type={'type': 'cosine_top', 'k': 2, 'proj_dim': 8, 'gate_noise': 1.0}, is_gshard_loss=False)
    x = torch.randn(4, 6, 16)
    y = layer(x)
    assert torch.isfinite(y).all() and torch.isfinite(y.l_aux)
    (y.sum() + y.l_aux).backward()
    assert layer.gates[0].sim_matrix.grad is not None


def test_custom_gate_and_expert_modules():
    class MyGate(torch.nn.Module):
        def __init__(self, model_dim, num_global_experts, k=1, **kw):
            super().__init__()
            self.top_k = k
            self.proj = torch.nn.Linear(model_dim, num_global_experts, bias=False)

        def forward(self, x):
            return self.proj(x)

    class MyExpert(torch.nn.Module):
        def __init__(self, model_dim, num_experts_per_device, sharded_count, scale=1.0):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(num_experts_per_device, 1, model_dim) * scale)

        def forward(self, x, ctx):
            assert ctx.num_local_experts == x.size(0)
            return x * self.w

    layer = moe.moe_layer(gate_type={'type': 'custom', 'module': MyGate, 'k': 1}, model_dim=8,
                          experts={'type': 'custom', 'module': MyExpert, 'num_experts_per_device': 3, 'scale': 2.0})
    x = torch.randn(10, 8)
    y = layer(x, capacity_factor=-100.0)      # negative: min(max count, 100 * samples/expert) => nothing dropped
    g = torch.softmax(layer.gates[0](x), 1).max(1)[0]
    assert torch.allclose(y, x * 2.0 * g.unsqueeze(1), atol=1e-6)
    assert all(hasattr(p, '_tutel_expert') for p in layer.experts.parameters())
    assert [n for n, _ in layer.get_parameter_iterator('gate')] == ['0.proj.weight']


def test_llama_ffn_expert_and_megablocks_flag_is_ignored_when_training():
    torch.manual_seed(0)
    layer = moe.moe_layer(gate_type={'type': 'top', 'k': 2}, model_dim=16,
                          experts={'type': 'llama_ffn', 'num_experts_per_device': 2, 'hidden_size_per_expert': 24})
    x = torch.randn(6, 16)
    y = layer(x, megablocks_size=2)
    assert layer.megablocks_size == 0 and y.shape == x.shape
    with torch.no_grad():
        y2 = make(E=4)(torch.randn(6, 16), megablocks_size=2, capacity_factor=-8.0)
    assert y2.shape == (6, 16)


def test_global_expert_count_rules():
    L = moe.moe_layer
    assert L.global_expert_count(3) == 3 and L.global_expert_count(-1) == 1
    with pytest.raises(Exception):
        L.global_expert_count(0)
