"""Merge per-rank MoE checkpoints into one file:  python -m tutel_b200.checkpoint.gather --inputs=./ckpt/{rank}-of-{size}.ckpt --input_size=8 --output=./model.ckpt

File-format parity with tutel/checkpoint/gather.py:12-78 (doc/CHECKPOINT.md): expert tensors are concatenated on
dim 0 in rank order; when several ranks share one expert (``input_size % E == 0``) the result is re-viewed to
``[E, -1, ...]`` so that hidden-dim slices re-join their expert.  Non-expert tensors are taken from rank 0.
"""
import argparse
import logging

import torch

from ..system import apply_rank_size_from_pattern
from .common import descend, expert_param_keys, legacy_prefixes, moe_layer_prefixes


def gather_states(states, default_num_global_experts=0):
    size = len(states)
    merged = dict(states[0])
    prefixes = moe_layer_prefixes(states[0])
    counts = {p: int(states[0][p + '_num_global_experts']) for p in prefixes}
    for p in legacy_prefixes(states[0]):
        if default_num_global_experts <= 0:
            raise Exception('Legacy checkpoint without `_num_global_experts`: please pass --default_num_global_experts')
        counts[p] = default_num_global_experts
        merged[p + '_num_global_experts'] = torch.tensor(default_num_global_experts)
    for prefix, E in counts.items():
        for key in expert_param_keys(states[0], prefix):
            joined = torch.cat([s[key] for s in states], dim=0)
            if size % E == 0 and size > E:
                joined = joined.view([E, -1] + list(joined.shape[2:])) if joined.dim() > 1 else joined.view(E, -1)
            merged[key] = joined
    return merged


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--inputs', type=str, required=True)
    ap.add_argument('--input_size', type=int, required=True)
    ap.add_argument('--output', type=str, required=True)
    ap.add_argument('--namespace', type=str, default='')
    ap.add_argument('--default_num_global_experts', type=int, default=0)
    args = ap.parse_args(argv)

    raw = [torch.load(apply_rank_size_from_pattern(args.inputs, rank=r, size=args.input_size, create_dir=False),
                      map_location='cpu') for r in range(args.input_size)]
    states = [descend(r, args.namespace) for r in raw]
    merged = gather_states(states, args.default_num_global_experts)
    out = raw[0]
    target = descend(out, args.namespace)
    target.clear()
    target.update(merged)
    torch.save(out, args.output)
    logging.warning('Gathered %d checkpoint shard(s) into %s', args.input_size, args.output)


if __name__ == '__main__':
    main()
