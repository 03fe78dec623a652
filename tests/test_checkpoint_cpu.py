"""Checkpoint re-sharding tools: format parity with doc/CHECKPOINT.md of the reference (gather / scatter round trips)."""
import os

import torch

from tutel_b200.checkpoint import gather, scatter


def _rank_state(rank, E_local=2, H=6, M=4):
    g = torch.Generator().manual_seed(rank)
    return {
        '_moe_layer._num_global_experts': torch.tensor(4),
        '_moe_layer.experts.batched_fc1_w': torch.randn(E_local, H, M, generator=g),
        '_moe_layer.experts.batched_fc2_w': torch.randn(E_local, H, M, generator=g),
        '_moe_layer.experts.batched_fc1_bias': torch.randn(E_local, H, generator=g),
        '_moe_layer.experts.batched_fc2_bias': torch.randn(E_local, M, generator=g),
        '_moe_layer.gates.0.wg.weight': torch.arange(16.).view(4, 4),
    }


def test_gather_then_scatter_roundtrip(tmp_path):
    for r in range(2):
        torch.save(_rank_state(r), tmp_path / ('%d-of-2.ckpt' % r))
    merged_path = str(tmp_path / 'merged.ckpt')
    gather.main(['--inputs', str(tmp_path / '{rank}-of-{size}.ckpt'), '--input_size', '2', '--output', merged_path])
    merged = torch.load(merged_path)
    assert merged['_moe_layer.experts.batched_fc1_w'].shape == (4, 6, 4)
    assert torch.equal(merged['_moe_layer.experts.batched_fc1_w'][2:], _rank_state(1)['_moe_layer.experts.batched_fc1_w'])
    # 4 experts -> 4 ranks (1 expert each) and -> 8 ranks (each expert split into 2 hidden slices)
    scatter.main(['--input', merged_path, '--output_size', '4', '--outputs', str(tmp_path / 'four/{rank}-of-{size}.ckpt')])
    one = torch.load(tmp_path / 'four' / '3-of-4.ckpt')
    assert torch.equal(one['_moe_layer.experts.batched_fc1_w'][0], merged['_moe_layer.experts.batched_fc1_w'][3])
    assert torch.equal(one['_moe_layer.gates.0.wg.weight'], merged['_moe_layer.gates.0.wg.weight'])
    scatter.main(['--input', merged_path, '--output_size', '8', '--outputs', str(tmp_path / 'eight/{rank}-of-{size}.ckpt')])
    half = torch.load(tmp_path / 'eight' / '5-of-8.ckpt')   # expert 2, second hidden slice
    assert half['_moe_layer.experts.batched_fc1_w'].shape == (1, 3, 4)
    assert torch.equal(half['_moe_layer.experts.batched_fc1_w'][0], merged['_moe_layer.experts.batched_fc1_w'][2, 3:])
    # and back: 8 shards -> one file identical to the merged one
    gather.main(['--inputs', str(tmp_path / 'eight/{rank}-of-{size}.ckpt'), '--input_size', '8', '--output', str(tmp_path / 'again.ckpt')])
    again = torch.load(tmp_path / 'again.ckpt')
    for k, v in merged.items():
        assert torch.equal(again[k].reshape(v.shape), v), k


def test_namespace_and_layer_loading(tmp_path):
    import torch.nn.functional as F
    from tutel_b200 import moe
    layer = moe.moe_layer(gate_type={'type': 'top', 'k': 1}, model_dim=4,
                          experts={'type': 'ffn', 'num_experts_per_device': 4, 'hidden_size_per_expert': 6,
                                   'activation_fn': lambda x: F.relu(x)})
    torch.save({'model': {'layer.' + k: v for k, v in layer.state_dict().items()}, 'epoch': 3}, tmp_path / 'full.ckpt')
    scatter.main(['--input', str(tmp_path / 'full.ckpt'), '--output_size', '2', '--outputs', str(tmp_path / 's/{rank}-{size}.ckpt'),
                  '--namespace', 'model'])
    part = torch.load(tmp_path / 's' / '1-2.ckpt')
    assert part['epoch'] == 3 and part['model']['layer.experts.batched_fc1_w'].shape == (2, 6, 4)
    assert torch.equal(part['model']['layer.experts.batched_fc1_w'], layer.experts.batched_fc1_w[2:])
