"""All-to-all / expert-FFN micro-pipelining (``a2a_ffn_overlap_degree``).

Reference: tutel/impls/overlap.py:8-67 + the stream/event autograd Functions of tutel/impls/communicate.py:257-397
+ the chunked NCCL all-to-alls of tutel/custom/custom_kernel.cpp:520-654.

The capacity dimension is cut into ``d`` chunks; the exchange of chunk *i+1* runs while the experts process chunk
*i*, in forward and - mirrored - in backward.  Two small autograd Functions express this without touching the
expert code: ``_Begin`` starts an exchange asynchronously (backward: waits for the mirrored exchange),
``_End`` waits for it (backward: starts the mirrored exchange).  Because autograd replays nodes in reverse creation
order, issuing all ``_Begin`` nodes first makes the backward pass pipeline itself the same way.

Results are bit-identical for every ``d`` (the experts are row-wise independent).  On the fused NVLink path
(:mod:`tutel_b200.parallel.fused`) the same knob selects the granularity of the arrival flags instead.
"""
from __future__ import annotations

from typing import Any, Callable, List

import torch
from . import communicate as C

MAX_OVERLAP_DEGREE = 32


class _Pending:
    """An in-flight exchange: the (not yet valid) output buffer and a callable that makes the current stream wait."""
    __slots__ = ('out', 'wait')

    def __init__(self):
        self.out, self.wait = None, None


_COMM_STREAMS = {}


def _comm_stream(device) -> torch.cuda.Stream:
    key = torch.device(device).index
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


def _start_exchange(packed: torch.Tensor, group, pending: _Pending, use_2dh: bool = False) -> None:
    """Start an all-to-all along dim 0 of ``packed`` (flat or 2-D hierarchical) without blocking the current stream."""
    packed = packed.contiguous()
    if not packed.is_cuda:
        pending.out, pending.wait = C._raw_all_to_all(packed, group, use_2dh), (lambda: None)
        return
    cur = torch.cuda.current_stream()
    side = _comm_stream(packed.device)
    side.wait_stream(cur)
    t = C._p2p(group, packed)
    with torch.cuda.stream(side):
        if t is not None and not use_2dh:
            # own counter slot: a main-stream P2P collective inside the expert function must not share epochs / mailboxes
            # with the exchanges that are in flight on this stream
            out = t.all_to_all(packed, slot=t.side_slot())
        else:
            out = C._raw_all_to_all(packed, group, use_2dh)      # NCCL, or the two phases of the hierarchical exchange
        ev = torch.cuda.Event()
        ev.record(side)
    packed.record_stream(side)
    out.record_stream(cur)
    # a result living in the transport's arena pool is only re-used by a later collective, which is ordered after
    # everything queued on the consumer stream (wait_stream above)
    pending.out, pending.wait = out, (lambda: torch.cuda.current_stream().wait_event(ev))


class _Begin(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, packed: torch.Tensor, group, fwd: _Pending, bwd: _Pending, use_2dh: bool):
        ctx.bwd = bwd
        _start_exchange(packed, group, fwd, use_2dh)
        return fwd.out

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):
        # the mirrored exchange was started by _End.backward; `grad` is its output buffer
        ctx.bwd.wait()
        return ctx.bwd.out, None, None, None, None


class _End(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, raw: torch.Tensor, group, fwd: _Pending, bwd: _Pending, use_2dh: bool):
        ctx.group, ctx.bwd, ctx.use_2dh = group, bwd, use_2dh
        fwd.wait()
        return raw.view_as(raw)

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):
        _start_exchange(grad, ctx.group, ctx.bwd, ctx.use_2dh)
        return ctx.bwd.out, None, None, None, None


def _async_all_to_all(x: torch.Tensor, input_dim: int, output_dim: int, group, use_2dh: bool = False) -> Callable[[], torch.Tensor]:
    """Start ``all_to_all(x, input_dim, output_dim)``; the returned callable finishes it (autograd-aware)."""
    world = C.get_world_size(group)
    fwd, bwd = _Pending(), _Pending()
    raw = _Begin.apply(C._a2a_pack(x, output_dim, world), group, fwd, bwd, use_2dh)
    return lambda: C._a2a_unpack(_End.apply(raw, group, fwd, bwd, use_2dh), input_dim)


def a2a_ffn_overlap_forward(input: torch.Tensor, expert_fn: Callable[[torch.Tensor], torch.Tensor],
                            a2a_ffn_overlap_degree: int, use_2dh: bool, group) -> torch.Tensor:
    """``all_to_all(1,0) -> experts -> all_to_all(0,1)`` on ``[E', C', M]`` pipelined over ``d`` capacity chunks."""
    d = a2a_ffn_overlap_degree
    assert d <= MAX_OVERLAP_DEGREE, 'Excepting a2a_ffn_overlap_degree (%d) <= %d.' % (d, MAX_OVERLAP_DEGREE)
    assert input.shape[1] % d == 0, 'Excepting input.shape[1] (%d) be multiple of a2a_ffn_overlap_degree (%d).' % (input.shape[1], d)
    if C.get_world_size(group) == 1:
        return expert_fn(input)
    chunks = input.split(input.shape[1] // d, dim=1)
    # (the hierarchical exchange pipelines the same way: both of its phases and the record transposes between them run on
    # the communication stream - the reference's AllToAll2DAsync, tutel/custom/custom_kernel.cpp:656-738)
    arrivals = [_async_all_to_all(c, 1, 0, group, use_2dh) for c in chunks]   # all dispatch exchanges are in flight
    returns: List[Callable[[], torch.Tensor]] = []
    for arrive in arrivals:
        y = expert_fn(arrive())                                           # waits only for its own chunk
        returns.append(_async_all_to_all(y, 0, 1, group, use_2dh))        # combine exchange overlaps the next chunk
    return torch.cat([r() for r in returns], dim=1)
