#!/usr/bin/env python3
"""Tiny launcher for profiling ONE GLU kernel under ncu:  python bench/glu_ncu_target.py {fwd|bwd} [act]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tutel_b200.ops import gemm as G

mode = sys.argv[1] if len(sys.argv) > 1 else 'bwd'
act = sys.argv[2] if len(sys.argv) > 2 else 'silu'
E, T, M, H = 8, 2048, 4096, 14336
torch.manual_seed(0)
x = (torch.randn(E, T, M, device='cuda') * 0.5).bfloat16()
w1 = (torch.randn(E, M, H, device='cuda') * 0.02).bfloat16()
w2 = (torch.randn(E, M, H, device='cuda') * 0.02).bfloat16()
w3 = (torch.randn(E, H, M, device='cuda') * 0.02).bfloat16()
dy = (torch.randn(E, T, M, device='cuda') * 0.5).bfloat16()
h, g, u = G.glu_gemm(x, w1, w2, b_mn=True, act=act, save_pre=True)
torch.cuda.synchronize()
for _ in range(2):
    if mode == 'fwd':
        G.glu_gemm(x, w1, w2, b_mn=True, act=act, save_pre=True)
    else:
        G.glu_gemm_bwd(dy, w3, g, u, b_mn=False, act=act)
torch.cuda.synchronize()
