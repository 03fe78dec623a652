"""MX block-scaled fp8 (OCP Microscaling: e4m3 elements, one UE8M0 power-of-two scale per 32 consecutive K elements).

The scales are consumed by the tensor core itself (``tcgen05.mma.kind::mxf8f6f4.block_scale``, csrc/gemm_mx.cu), so the
finer granularity costs no epilogue work.  The reference has no reduced-precision expert path (its experts run
``torch.matmul`` in the model dtype, tutel/experts/ffn.py); the framework's fused engine uses the row-scaled e4m3 GEMM of
csrc/gemm_sm100.cu, this module is the finer-grained alternative for GEMMs whose rows carry outliers.

Everything here also has a pure PyTorch definition (``*_reference``) that runs on CPU: the tests compare the kernels
against it, and it documents the number format:

* block exponent   ``e = ceil(log2(amax / 448))`` clamped to [-127, 126]
* elements         ``q = e4m3_rn(x * 2**-e)``
* scale byte       ``e + 127``
* scale storage    ``sf[g][k // 128][r // 128][(r % 32) * 16 + ((r % 128) // 32) * 4 + (k % 128) // 32]`` - 512-byte
  atoms in the order the tensor core reads them from tensor memory (rows padded to a multiple of 128 with byte 0).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import backend

BLOCK = 32
E4M3_MAX = 448.0


def _check_k(K: int):
    if K % 128 != 0:
        raise ValueError('MX operands need K %% 128 == 0 (got %d)' % K)


def block_exponents_reference(x: torch.Tensor) -> torch.Tensor:
    """Shared exponents [.., K / 32] (int32) of the 32-element blocks along the last dim."""
    amax = x.float().abs().reshape(*x.shape[:-1], x.shape[-1] // BLOCK, BLOCK).amax(-1)
    # ceil(log2(v)) from the float representation: exponent field (+1 when the mantissa is non-zero), as the kernel does
    v = (amax * (1.0 / E4M3_MAX)).contiguous()
    bits = v.view(torch.int32)
    e = ((bits >> 23) & 0xFF) - 127 + ((bits & 0x7FFFFF) != 0).to(torch.int32)
    return e.clamp_(-127, 126)


def pack_scales(e: torch.Tensor) -> torch.Tensor:
    """Block exponents [G, R, K / 32] -> scale bytes in tile order (uint8, flat)."""
    G, R, KB32 = e.shape
    K = KB32 * BLOCK
    _check_k(K)
    RT = (R + 127) // 128
    b = torch.zeros(G, RT * 128, KB32, dtype=torch.uint8, device=e.device)
    b[:, :R] = (e + 127).to(torch.uint8)
    # [G, rt, c(4), l(32), kb, j(4)] -> [G, kb, rt, l, c, j]
    b = b.view(G, RT, 4, 32, K // 128, 4).permute(0, 4, 1, 3, 2, 5)
    return b.contiguous().view(-1)


def unpack_scales(sf: torch.Tensor, G: int, R: int, K: int) -> torch.Tensor:
    """Inverse of :func:`pack_scales`: block exponents [G, R, K / 32] (int32)."""
    _check_k(K)
    RT = (R + 127) // 128
    b = sf.view(G, K // 128, RT, 32, 4, 4).permute(0, 2, 4, 3, 1, 5).contiguous().view(G, RT * 128, K // BLOCK)
    return b[:, :R].to(torch.int32) - 127


def mx_quantize_reference(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [G, R, K] -> (q e4m3 [G, R, K], sf uint8) in pure PyTorch (any device)."""
    G, R, K = x.shape
    _check_k(K)
    e = block_exponents_reference(x)
    inv = torch.exp2(-e.float()).unsqueeze(-1)
    q = (x.float().view(G, R, K // BLOCK, BLOCK) * inv).clamp_(-E4M3_MAX, E4M3_MAX).view(G, R, K).to(torch.float8_e4m3fn)
    return q, pack_scales(e)


def mx_dequantize(q: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    """fp32 values of an MX operand (pure PyTorch)."""
    G, R, K = q.shape
    e = unpack_scales(sf, G, R, K)
    return (q.float().view(G, R, K // BLOCK, BLOCK) * torch.exp2(e.float()).unsqueeze(-1)).view(G, R, K)


def mx_quantize(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [G, R, K] (or [R, K]) bf16 / fp16 -> (q, sf).  One launch of ``mx_quantize_kernel`` on a GPU."""
    squeeze = x.dim() == 2
    x3 = x.unsqueeze(0) if squeeze else x
    _check_k(x3.shape[-1])
    if x3.is_cuda and x3.element_size() == 2 and backend.has_ext():
        backend.count_launch()
        q, sf = backend.require_ext().mx_quantize(x3.contiguous())
    else:
        if x3.is_cuda and not backend.allow_fallback():
            raise RuntimeError('mx_quantize: 16-bit CUDA input and the native extension are required on a GPU')
        q, sf = mx_quantize_reference(x3)
    return (q[0] if squeeze else q), sf


def mx_gemm(a: torch.Tensor, sfa: torch.Tensor, b: torch.Tensor, sfb: torch.Tensor, relu: bool = False,
            block_n: int = 0, _sf_addr_plain: bool = False) -> torch.Tensor:
    """``a [G, M, K] @ b [G, N, K]^T`` -> bf16 [G, M, N]; both operands from :func:`mx_quantize`."""
    if a.is_cuda and backend.has_ext():
        backend.count_launch()
        return backend.require_ext().mx_gemm(a, sfa, b, sfb, bool(relu), int(block_n), bool(_sf_addr_plain))
    if a.is_cuda and not backend.allow_fallback():
        raise RuntimeError('mx_gemm: the native extension is required on a GPU')
    y = torch.matmul(mx_dequantize(a, sfa), mx_dequantize(b, sfb).transpose(1, 2))
    return (torch.relu(y) if relu else y).to(torch.bfloat16)


_MX_WEIGHT_CACHE = {}


def mx_weight(w: torch.Tensor, transpose: bool = False):
    """MX copy of a weight [G, N, K] (or of ``w^T`` when ``transpose``), cached like ``ops.gemm.fp8_operand``: valid until
    the next optimizer step or in-place modification."""
    import weakref
    from . import gemm as _gemm
    _gemm._ensure_step_hook()
    anchor = w._base if w._base is not None else w
    key = (id(anchor), w.data_ptr(), bool(transpose), tuple(w.shape), tuple(w.stride()))
    stamp = (w._version, _gemm._FP8_STEP[0])
    hit = _MX_WEIGHT_CACHE.get(key)
    if hit is not None and hit[0] == stamp and hit[3]() is anchor:
        return hit[1], hit[2]
    src = w.detach()
    q, sf = mx_quantize((src.transpose(1, 2) if transpose else src).contiguous())
    if len(_MX_WEIGHT_CACHE) > 256:
        for k in [k for k, v in _MX_WEIGHT_CACHE.items() if v[3]() is None]:
            del _MX_WEIGHT_CACHE[k]
    _MX_WEIGHT_CACHE[key] = (stamp, q, sf, weakref.ref(anchor))
    return q, sf


def mx_linear(x: torch.Tensor, w: torch.Tensor, w_layout: str = 'nk', relu: bool = False) -> torch.Tensor:
    """``x [G, R, K] @ W`` with both operands quantised to MX fp8 on the fly (weights cached); W is [G, N, K] ('nk') or
    [G, K, N] ('kn').  Inference / forward only."""
    xq, xs = mx_quantize(x)
    wq, ws = mx_weight(w, transpose=(w_layout == 'kn'))
    return mx_gemm(xq, xs, wq, ws, relu=relu)


def mx_ffn(x: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """Expert FFN forward ``relu(x @ w1^T) @ w2`` for x [E, C, M], w1 [E, H, M], w2 [E, H, M] (the layout of
    ``models/experts/ffn.py``) in MX fp8: 2 quantisation launches + 2 GEMMs (ReLU fused into the first)."""
    h = mx_linear(x, w1, 'nk', relu=True)
    return mx_linear(h, w2, 'kn')
