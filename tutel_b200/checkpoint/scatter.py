"""Split one merged MoE checkpoint into per-rank files:  python -m tutel_b200.checkpoint.scatter --input=./model.ckpt --output_size=2 --outputs=./new/{rank}-of-{size}.ckpt

File-format parity with tutel/checkpoint/scatter.py:11-72: with ``E % N == 0`` every rank receives ``E/N`` whole
experts; with ``N % E == 0`` each expert is flat-split into ``N/E`` contiguous pieces (== hidden-dimension slices of
the ``[H, M]`` matrices), which is exactly the layout the layer expects for ``num_experts_per_device = -N/E``.
"""
import argparse
import logging

import torch

from ..system import apply_rank_size_from_pattern
from .common import descend, expert_param_keys, moe_layer_prefixes


def scatter_state(state, size):
    shards = [dict(state) for _ in range(size)]
    for prefix in moe_layer_prefixes(state):
        E = int(state[prefix + '_num_global_experts'])
        for key in expert_param_keys(state, prefix):
            full = state[key]
            if E % size == 0:
                pieces = full.view([size, E // size] + list(full.shape[1:])).unbind(0)
            elif size % E == 0:
                per = size // E
                flat = full.reshape(E, per, -1)
                pieces = []
                for r in range(size):
                    piece = flat[r // per, r % per]
                    if full.dim() > 2:
                        piece = piece.view([1, -1] + list(full.shape[2:]))
                    else:
                        piece = piece.view(1, -1)
                    pieces.append(piece)
            else:
                raise Exception('Cannot scatter %d experts onto %d devices' % (E, size))
            for r in range(size):
                shards[r][key] = pieces[r].contiguous().clone()
    return shards


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--input', type=str, required=True)
    ap.add_argument('--output_size', type=int, required=True)
    ap.add_argument('--outputs', type=str, required=True)
    ap.add_argument('--namespace', type=str, default='')
    args = ap.parse_args(argv)

    raw = torch.load(args.input, map_location='cpu')
    shards = scatter_state(descend(raw, args.namespace), args.output_size)
    for r, shard in enumerate(shards):
        target = descend(raw, args.namespace)
        backup = dict(target)
        target.clear()
        target.update(shard)
        torch.save(raw, apply_rank_size_from_pattern(args.outputs, rank=r, size=args.output_size))
        target.clear()
        target.update(backup)
    logging.warning('Scattered %s into %d shard(s)', args.input, args.output_size)


if __name__ == '__main__':
    main()
