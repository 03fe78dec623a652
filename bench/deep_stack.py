#!/usr/bin/env python3
"""A deep stack of flagship-shape MoE layers (top-2, 8 global experts, M=4096, H=14336, bf16, 8192 tokens/GPU) trained
through the fused engine inside a SMALL symmetric heap: all layers share one ring of three (IN, OUT) buffer sets, leases
are spilled when the ring wraps (parallel/fused.py).

    TUTEL_B200_HEAP_MB=4096 TUTEL_B200_STAGE_MB=1024 torchrun --nproc-per-node=8 bench/deep_stack.py --layers 16
Prints one JSON line (rank 0): ms/step (device-timed, max over ranks), arena use, whether every layer ran fused."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TUTEL_B200_HEAP_MB', '4096')
os.environ.setdefault('TUTEL_B200_STAGE_MB', '1024')

import torch
import torch.distributed as dist
import torch.nn.functional as F

from tutel_b200 import moe, net, system
from tutel_b200.parallel import fused, p2p

ap = argparse.ArgumentParser()
ap.add_argument('--layers', type=int, default=16)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--model_dim', type=int, default=4096)
ap.add_argument('--hidden', type=int, default=14336)
ap.add_argument('--tokens', type=int, default=8192)
args = ap.parse_args()

env = system.init_data_model_parallel(backend='nccl')
rank, world, dev = env.global_rank, env.global_size, env.local_device
torch.set_default_dtype(torch.bfloat16)
layers = torch.nn.ModuleList([
    moe.moe_layer(gate_type={'type': 'top', 'k': 2, 'capacity_factor': 1.0}, model_dim=args.model_dim,
                  experts={'type': 'ffn', 'num_experts_per_device': 8 // world, 'hidden_size_per_expert': args.hidden,
                           'activation_fn': lambda t: F.relu(t)},
                  scan_expert_func=lambda n, p: setattr(p, 'skip_allreduce', True), seeds=(1, rank + 1, 1)) for _ in range(args.layers)]).to(dev)
opt = torch.optim.SGD(layers.parameters(), lr=1e-6)
shared = [p for p in layers.parameters() if not hasattr(p, 'skip_allreduce')]
torch.manual_seed(rank)
x = torch.randn(args.tokens, args.model_dim, device=dev)
calls = {'fused': 0, 'generic': 0}
orig = fused.engine_for


def counting(layer, xx, crit, d):
    r = orig(layer, xx, crit, d)
    calls['fused' if r is not None else 'generic'] += 1
    return r


fused.engine_for = counting


def step():
    opt.zero_grad()
    h = x
    for layer in layers:
        h = h + layer(h)
    loss = h.float().pow(2).mean()
    loss.backward()
    for p in shared:
        p.grad = net.simple_all_reduce(p.grad) / world
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
t = p2p.transport_for(None)
if rank == 0:
    rings = [r for r in fused._engine(t).rings.values() if r is not None] if t is not None else []
    print(json.dumps({'layers': args.layers, 'world': world, 'ms_per_step': float(ms), 'loss': float(loss),
                      'heap_mb': t.heap_bytes >> 20 if t else None, 'heap_used_mb': (t._bump >> 20) if t else None,
                      'ring_sets': [len(r.sets) for r in rings], 'fused_calls': calls['fused'], 'generic_calls': calls['generic'],
                      'gpu_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}))
