#!/usr/bin/env python3
"""BASELINE config #4: dropless (capacity_factor=0) decoder-style inference, 128 local experts on one GPU,
32 tokens, top-1, model_dim = hidden = 2048, fp32 - the reference's "Megablocks" demo (README.md:52-58).

    python bench/dropless_bench.py --impl ours      --megablocks_size 1
    python bench/dropless_bench.py --impl reference --megablocks_size 1
Device-timed forward latency (CUDA events, L2 flushed between iterations), one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
ap.add_argument('--megablocks_size', type=int, default=1)
ap.add_argument('--experts', type=int, default=128)
ap.add_argument('--tokens', type=int, default=32)
ap.add_argument('--dim', type=int, default=2048)
ap.add_argument('--dtype', default='float32')
ap.add_argument('--iters', type=int, default=50)
ap.add_argument('--graph', action='store_true', help='ours only: replay the forward as one CUDA graph (tutel_b200.utils.graph)')
args = ap.parse_args()
if args.impl == 'reference':
    sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
    from tutel import moe, system
else:
    sys.path.insert(0, ROOT)
    from tutel_b200 import moe, system
import torch
import torch.nn.functional as F

env = system.init_data_model_parallel(backend='nccl')
dev = env.local_device
torch.set_default_dtype(getattr(torch, args.dtype))
torch.manual_seed(0)
layer = moe.moe_layer(gate_type={'type': 'top', 'k': 1, 'capacity_factor': 0.0}, model_dim=args.dim,
                      experts={'type': 'ffn', 'num_experts_per_device': args.experts, 'hidden_size_per_expert': args.dim,
                               'activation_fn': lambda x: F.relu(x)}, seeds=(1, 1, 1)).to(dev).eval()
x = torch.randn(1, args.tokens, args.dim, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
call = (lambda t: layer(t, megablocks_size=args.megablocks_size)) if args.megablocks_size > 0 else (lambda t: layer(t))
if args.graph and args.impl == 'ours':
    from tutel_b200.utils.graph import GraphedForward
    call = GraphedForward(call, x)
times = []
with torch.no_grad():
    for i in range(args.iters + 5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = call(x)
        e.record()
        torch.cuda.synchronize()
        if i >= 5:
            times.append(s.elapsed_time(e))
times.sort()
print(json.dumps({'impl': args.impl, 'config': 'dropless cf=0 top-1 E=%d tokens=%d dim=%d %s megablocks_size=%d%s' % (
    args.experts, args.tokens, args.dim, args.dtype, args.megablocks_size, ' cuda-graph' if args.graph and args.impl == 'ours' else ''), 'median_ms': times[len(times) // 2], 'min_ms': times[0],
    'checksum': float(y.float().abs().sum())}))
