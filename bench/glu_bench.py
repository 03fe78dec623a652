#!/usr/bin/env python3
"""Isolated timings of the SwiGLU-expert GEMMs (Mixtral shape) - ours vs cuBLAS + eager elementwise.
    python bench/glu_bench.py [--experts 8 --rows 2048 --model_dim 4096 --hidden 14336]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from tutel_b200.ops import gemm as G

ap = argparse.ArgumentParser()
ap.add_argument('--experts', type=int, default=8)
ap.add_argument('--rows', type=int, default=2048)
ap.add_argument('--model_dim', type=int, default=4096)
ap.add_argument('--hidden', type=int, default=14336)
ap.add_argument('--act', type=str, default='silu')
ap.add_argument('--iters', type=int, default=8)
ap.add_argument('--json', type=str, default='')
args = ap.parse_args()
E, T, M, H = args.experts, args.rows, args.model_dim, args.hidden
dev = 'cuda'
torch.manual_seed(0)
x = (torch.randn(E, T, M, device=dev) * 0.5).bfloat16()
w1 = (torch.randn(E, M, H, device=dev) * 0.02).bfloat16()
w2 = (torch.randn(E, M, H, device=dev) * 0.02).bfloat16()
w3 = (torch.randn(E, H, M, device=dev) * 0.02).bfloat16()
dy = (torch.randn(E, T, M, device=dev) * 0.5).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
act_fn = {'silu': F.silu, 'relu': torch.relu, 'gelu': F.gelu}[args.act]


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


unit = 2.0 * E * T * M * H * 1e-9        # GFLOP of one [T,M]x[M,H] product over all experts
h, g, u = G.glu_gemm(x, w1, w2, b_mn=True, act=args.act, save_pre=True)
dg, du = G.glu_gemm_bwd(dy, w3, g, u, b_mn=False, act=args.act)
res = {}


def rec(name, ms, units):
    res[name] = {'ms': ms, 'tflops': units * unit / ms}
    print('%-44s %8.3f ms  %7.1f TFLOP/s' % (name, ms, units * unit / ms), flush=True)


rec('glu fwd (dual-B, saves g,u,h)', timeit(lambda: G.glu_gemm(x, w1, w2, b_mn=True, act=args.act, save_pre=True)), 2)
rec('glu fwd (inference, h only)', timeit(lambda: G.glu_gemm(x, w1, w2, b_mn=True, act=args.act)), 2)
rec('eager: 2 cuBLAS GEMMs + act + mul', timeit(lambda: act_fn(torch.matmul(x, w1)) * torch.matmul(x, w2)), 2)
rec('down proj h@W3 (ours)', timeit(lambda: G.raw_gemm(h, w3, b_mn=True)), 1)
rec('down proj h@W3 (cuBLAS)', timeit(lambda: torch.matmul(h, w3)), 1)
rec('glu bwd dh GEMM -> dg,du (ours)', timeit(lambda: G.glu_gemm_bwd(dy, w3, g, u, b_mn=False, act=args.act)), 1)


def eager_bwd():
    dh = torch.matmul(dy, w3.transpose(1, 2))
    gf = g.float()
    if args.act == 'relu':
        a, da = torch.relu(g), (g > 0).to(g.dtype)
    else:
        sg = torch.sigmoid(gf)
        a, da = (gf * sg).to(g.dtype), (sg * (1 + gf * (1 - sg))).to(g.dtype)
    return dh * u * da, dh * a


rec('eager: cuBLAS dh + elementwise dg,du', timeit(eager_bwd), 1)
rec('wgrad dW1 = x^T@dg (ours)', timeit(lambda: G.raw_gemm(x, dg, a_mn=True, b_mn=True)), 1)
rec('wgrad dW1 = x^T@dg (cuBLAS)', timeit(lambda: torch.matmul(x.transpose(1, 2), dg)), 1)
rec('wgrad dW3 = h^T@dy (ours)', timeit(lambda: G.raw_gemm(h, dy, a_mn=True, b_mn=True)), 1)
rec('wgrad dW3 = h^T@dy (cuBLAS)', timeit(lambda: torch.matmul(h.transpose(1, 2), dy)), 1)
rec('dgrad dx = dg@W1^T (+ fused add) (ours)', timeit(lambda: G.raw_gemm(du, w2, epilogue=G.EPI_ADD, aux=G.raw_gemm(dg, w1))), 2)
if args.json:
    os.makedirs(os.path.dirname(args.json) or '.', exist_ok=True)
    with open(args.json, 'w') as f:
        json.dump({'config': vars(args), 'results': res}, f, indent=1)
