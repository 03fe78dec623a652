"""Location scan helper kept for API parity (``tutel.moe.fast_cumsum_sub_one``, tutel/jit_kernels/gating.py:19-24).

The layer itself no longer scans one-hot masks: routing uses the fused histogram/scan/rank kernels of
csrc/moe_kernels.cu (see :mod:`tutel_b200.ops.routing`).
"""
import torch


def fast_cumsum_sub_one(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    if data.dim() != 2 or dim != 0:
        raise Exception('Unimplemented fast_cumsum_sub_one() of data = %s and dim = %s' % (data.size(), dim))
    return torch.cumsum(data, dim=0) - 1
