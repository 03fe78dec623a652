"""Location scan helper kept for API parity (``tutel.moe.fast_cumsum_sub_one``, tutel/jit_kernels/gating.py:19-24).

The layer itself no longer scans one-hot masks: routing uses the fused histogram/scan/rank kernels of
csrc/moe_kernels.cu (see :mod:`tutel_b200.ops.routing`).
"""
import torch


def fast_cumsum_sub_one(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    if data.dim() != 2 or dim != 0:
        raise Exception('Unimplemented fast_cumsum_sub_one() of data = %s and dim = %s' % (data.size(), dim))
    return torch.cumsum(data, dim=0) - 1


# ----------------------------------------------------------------------------------------------------------------
# Fused top-k gating (opt-in: TUTEL_B200_FUSED_GATE=1)
# ----------------------------------------------------------------------------------------------------------------
# The reference computes softmax, top-k, the one-hot masks, the GShard loss and the gate normalisation as ~15 separate
# PyTorch kernels per forward (tutel/impls/moe_layer.py:283-305, fast_dispatch.py:143-176, losses.py:12-19), and
# autograd adds as many again in backward.  Here: ONE kernel forward (one warp per token: softmax in registers,
# iterative arg-max, per-block partial sums for the loss) and ONE kernel backward (closed-form gradient of the
# normalised gates and of the loss through the softmax).  The pure-torch branch implements the same formulas and is
# what the CPU tests check against autograd of the unfused path.
def fused_gate_enabled() -> bool:
    import os
    return os.environ.get('TUTEL_B200_FUSED_GATE', '0') not in ('0', '', 'off', 'false')


class FusedTopKGate(torch.autograd.Function):
    """``logits [S,E] -> (gates [k,S], l_aux)`` plus non-differentiable ``idx [k,S] int32``, ``top1 [S]`` (raw best score)."""

    @staticmethod
    def forward(ctx, logits: torch.Tensor, k: int, normalize: bool, want_loss: bool):
        from . import backend
        S, E = logits.shape
        eps = float(torch.finfo(logits.dtype).eps)
        lf = logits.detach().to(torch.float32).contiguous()
        use_kernel = lf.is_cuda and backend.has_cuda_ext() and E <= 512
        if use_kernel:
            backend.count_launch()
            p, idx, top, me_part, ce_part = backend.require_ext().gate_topk_forward(lf, k)
            me = me_part.sum(0)
            ce = ce_part.sum(0).to(torch.float32)
        else:
            p = torch.softmax(lf, dim=1)
            top_sk, idx_sk = torch.topk(p, k, dim=1)
            idx, top = idx_sk.t().contiguous().to(torch.int32), top_sk.t().contiguous()
            me = p.sum(0)
            ce = torch.zeros([E], dtype=torch.float32, device=p.device)
            ce.scatter_add_(0, idx[0].to(torch.int64), torch.ones([S], dtype=torch.float32, device=p.device))
        gates = top
        if normalize and k > 1:
            gates = top / torch.clamp(top.sum(dim=0, keepdim=True), min=eps)
        l_aux = (me * ce).sum() * (E / float(S * S)) if want_loss else None
        ctx.save_for_backward(p, idx, top, ce)
        ctx.k, ctx.normalize, ctx.eps, ctx.use_kernel, ctx.in_dtype, ctx.want_loss = k, normalize, eps, use_kernel, logits.dtype, want_loss
        out_loss = l_aux.to(logits.dtype) if want_loss else torch.zeros((), dtype=logits.dtype, device=logits.device)
        idx_out, top1 = idx, top[0].to(logits.dtype)
        ctx.mark_non_differentiable(idx_out, top1)
        return gates.to(logits.dtype), out_loss, idx_out, top1

    @staticmethod
    def backward(ctx, dgates, dloss, _didx, _dtop1):
        p, idx, top, ce = ctx.saved_tensors
        S, E = p.shape
        k = ctx.k
        dg = (dgates if dgates is not None else torch.zeros_like(top)).to(torch.float32).contiguous()
        dl = dloss.to(torch.float32).reshape(1).contiguous() if (ctx.want_loss and dloss is not None) else None
        if ctx.use_kernel:
            from . import backend
            backend.count_launch()
            dlogits = backend.require_ext().gate_topk_backward(p, idx, top, dg, ce if dl is not None else None, dl,
                                                               bool(ctx.normalize), ctx.eps)
        else:
            dr = dg
            if ctx.normalize and k > 1:
                D = top.sum(dim=0, keepdim=True)
                Dc = torch.clamp(D, min=ctx.eps)
                dot = (dg * top).sum(dim=0, keepdim=True)
                dr = dg / Dc - torch.where(D > ctx.eps, dot / (Dc * Dc), torch.zeros_like(dot))
            dp = torch.zeros_like(p)
            if dl is not None:
                dp += (dl * (E / float(S * S))) * ce.unsqueeze(0)
            dp.scatter_add_(1, idx.t().to(torch.int64), dr.t().contiguous())
            dlogits = p * (dp - (dp * p).sum(dim=1, keepdim=True))
        return dlogits.to(ctx.in_dtype), None, None, None


def fused_topk_gate(logits: torch.Tensor, k: int, normalize: bool = True, want_loss: bool = True):
    """Returns ``(idx_ks int32 [k,S], gates_ks [k,S], l_aux or None, top1 [S])``."""
    gates, l_aux, idx, top1 = FusedTopKGate.apply(logits, int(k), bool(normalize), bool(want_loss))
    return idx, gates, (l_aux if want_loss else None), top1
