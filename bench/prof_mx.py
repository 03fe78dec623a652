#!/usr/bin/env python3
"""Tiny driver for ncu: the MX block-scaled GEMM (csrc/gemm_mx.cu) and, for comparison, the row-scaled fp8 GEMM of
csrc/gemm_sm100.cu at the flagship expert shape D[16384, 14336] = A[16384, 4096] . B[14336, 4096]^T.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/mx_ncu python bench/prof_mx.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tutel_b200.ops import gemm, mx

M, N, K = 16384, 14336, 4096
a = (torch.randn(1, M, K, device='cuda') * 0.5).bfloat16()
b = (torch.randn(1, N, K, device='cuda') * 0.5).bfloat16()
aq, sa = mx.mx_quantize(a)
bq, sb = mx.mx_quantize(b)
rq, rs = gemm.quantize_rows(a)
wq, ws = gemm.quantize_rows(b)
d = torch.empty(1, M, N, device='cuda', dtype=torch.bfloat16)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    if which in ('all', 'cg2'):
        mx.mx_gemm(aq, sa, bq, sb, cta_group=2)
    if which in ('all', 'cg1'):
        mx.mx_gemm(aq, sa, bq, sb, cta_group=1)
    if which in ('all', 'row'):
        gemm.raw_gemm(rq, wq, out=d, scale_a=rs, scale_b=ws)
    if which in ('all', 'quant'):
        mx.mx_quantize(a)
        mx.mx_quantize_transpose(b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('prof_mx done')
