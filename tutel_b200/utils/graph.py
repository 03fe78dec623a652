"""CUDA-graph capture of launch-bound inference calls.

Small-batch (decoder) MoE inference is a dozen kernels of a few microseconds each: issued one by one from Python the
GPU idles between them and the layer costs whatever the CPU needs to launch it.  With a positive capacity factor, or
with the bound-based dropless mode (``capacity_factor <= 0`` + ``megablocks_size > 0`` on one GPU, see
models/moe_layer.py), a forward pass never touches the host, so the whole call can be recorded once and replayed as
ONE graph launch:

    fast = GraphedForward(lambda x: layer(x, megablocks_size=1), example_x)
    y = fast(x)                      # copies x into the static input, replays, returns the static output

The reference cannot do this: its dropless path reads the capacity and the per-expert counts back to the host
(tutel/impls/fast_dispatch.py:192-193, tutel/custom/custom_kernel.cpp:875).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedForward:
    """Record ``fn(*inputs)`` (no autograd) into a CUDA graph; calls replay it on fresh input values of the same shapes."""

    def __init__(self, fn: Callable, *example_inputs: torch.Tensor, warmup: int = 3):
        assert all(t.is_cuda for t in example_inputs), 'GraphedForward needs CUDA tensors'
        self._inputs = [t.detach().clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):       # lazy initialisation (kernel attributes, workspaces, caches) happens here
                fn(*self._inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self._graph):
            self._outputs = fn(*self._inputs)

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self._inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._outputs

    @property
    def static_inputs(self) -> Sequence[torch.Tensor]:
        return self._inputs
