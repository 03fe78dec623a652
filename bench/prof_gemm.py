#!/usr/bin/env python3
"""Tiny driver for ncu: flagship expert GEMM on the tcgen05 kernel (cta_group 1 and 2) and on cuBLAS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tutel_b200 import _C

M, N, K = 16384, 14336, 4096
a = (torch.randn(1, M, K, device='cuda') * 0.5).bfloat16()
b = (torch.randn(1, N, K, device='cuda') * 0.5).bfloat16()
d = torch.empty(1, M, N, device='cuda', dtype=torch.bfloat16)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for it in range(3):
    if it == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    if which in ('all', 'cg2'):
        _C.gemm(a, b, d, False, False, 0, None, None, None, 1.0, 1, 2, 256, 0, 0, 0, 0, 0, 0, 0, 0, 1, None, None, None)
    if which in ('all', 'cg1'):
        _C.gemm(a, b, d, False, False, 0, None, None, None, 1.0, 1, 1, 256, 0, 0, 0, 0, 0, 0, 0, 0, 1, None, None, None)
    if which in ('all', 'cublas'):
        torch.matmul(a, b.transpose(1, 2), out=d)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
