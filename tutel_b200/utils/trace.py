"""Optional NVTX ranges around the stages of the MoE layer (``TUTEL_B200_NVTX=1``).

The reference has no tracing hooks (only ``system.record_time`` and the SKIP_* ablation switches); here the
stages (route / encode / dispatch / experts / combine / decode, or ``fused`` for the single-engine path) show up as named
ranges in Nsight Systems / ``ncu --nvtx`` so a timeline can be read without guessing from kernel names.  Disabled
(the default) it costs one dictionary lookup per stage.
"""
from __future__ import annotations

import contextlib
import os

import torch

_NULL = contextlib.nullcontext()
_ENABLED = None


def enabled() -> bool:
    global _ENABLED
    if _ENABLED is None:
        _ENABLED = os.environ.get('TUTEL_B200_NVTX', '0') not in ('0', '', 'off', 'false') and torch.cuda.is_available()
    return _ENABLED


class _Range:
    __slots__ = ('name',)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


def stage(name: str):
    """``with stage('moe.encode'): ...``"""
    return _Range('tutel_b200.' + name) if enabled() else _NULL
