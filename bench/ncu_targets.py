#!/usr/bin/env python3
"""Launch every hot single-GPU kernel once at the flagship shape, for ncu (one GPU only):

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels python bench/ncu_targets.py
The profiled region (cudaProfilerStart/Stop) contains, in order: gate_route + route_finish, encode (local), encode fp8,
GEMM fwd (bias+relu), GEMM2 (b_mn), dgrad (relu-grad + colsum), wgrad (a_mn, b_mn), decode, gate_grad, colsum,
gate_route_bwd, skinny_ffn."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tutel_b200 import _C
from tutel_b200.ops import gemm as G

dev = 'cuda'
torch.manual_seed(0)
S, E, k, M, H = 8192, 8, 2, 4096, 14336
C = k * S // E
logits = torch.randn(S, E, device=dev).bfloat16()
x = torch.randn(S, M, device=dev).bfloat16()
w1 = (torch.randn(E, H, M, device=dev) * 0.02).bfloat16()
w2 = (torch.randn(E, H, M, device=dev) * 0.02).bfloat16()
b1 = torch.zeros(E, H, device=dev).bfloat16()
xs = torch.randn(128, 4, 2048, device=dev)
sw1, sw2 = torch.randn(128, 2048, 2048, device=dev) * 0.02, torch.randn(128, 2048, 2048, device=dev) * 0.02
cnt = (torch.rand(128, device=dev) < 0.22).int()


def run():
    scores, idx, top, gates, loc, counts, ce, l_aux, slot = _C.gate_route_forward(logits, k, C, True, 1e-3)
    buf = torch.empty(E * C, M, device=dev, dtype=torch.bfloat16)
    _C.encode_rows(x, None, slot, buf, k, E, C, 0, 0, 0, 0, 0, 0, None)
    q, sc = _C.encode_rows_fp8(x, None, slot, k, E, C, 0, 0, 0, 0, 0, 0, 0)
    act = G.raw_gemm(buf.view(E, C, M), w1, epilogue=G.EPI_BIAS_RELU, bias=b1)
    y = G.raw_gemm(act, w2, b_mn=True)
    db = torch.zeros(E, H, device=dev)
    dh = G.raw_gemm(y, w2, epilogue=G.EPI_RELU_BWD, aux=act, colsum=db)
    dw = G.raw_gemm(dh, buf.view(E, C, M), a_mn=True, b_mn=True)
    out = _C.decode_rows(y.view(E * C, M), gates, idx, loc, E, C, 0, 0)
    dg = _C.gate_grad(out, y.view(E * C, M), idx, loc, E, C)
    cs = _C.grouped_colsum(y)
    dlog = _C.gate_route_backward(scores, idx, top, dg, ce, l_aux, logits, True, 1e-3)
    ys = _C.skinny_ffn(xs, sw1, None, sw2, None, cnt, 1)
    return dlog, cs, dw, ys, q


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('ncu targets done')
