"""ZeRO-1/2 style optimizer wrapper with the interface of ``tutel.net.TutelDistributedOptimizer`` (behaviour:
tutel/net.py:15-58), organised around flat buckets instead of per-parameter collectives.

All dense (non-expert) parameters of one dtype live in ONE flat, padded buffer; the parameters and their ``.grad``
tensors are views into it.  A step is then, per bucket: one reduce-scatter of the flat gradient, the wrapped optimizer
stepping this rank's 1/W shard (its state - momentum, Adam moments - exists for that shard only), one all-gather of the
updated shard straight into the flat parameter buffer.  Nothing is copied per parameter and the number of collectives
does not grow with the number of layers (the reference issues two collectives and three small copies per parameter).
Expert parameters (tagged ``_tutel_expert`` by the MoE layer) are owned by one rank and stepped locally.
"""
from typing import Dict, List

import torch

from . import communicate as C


class _Bucket:
    """Flat storage of same-dtype parameters: ``flat`` / ``grad`` are [W * shard_len]; this rank owns one shard."""

    def __init__(self, params: List[torch.nn.Parameter], world: int, rank: int):
        self.params = params
        numel = sum(p.numel() for p in params)
        self.shard_len = (numel + world - 1) // world
        dev, dt = params[0].device, params[0].dtype
        self.flat = torch.zeros([world * self.shard_len], dtype=dt, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.views = []
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                g = self.grad[off:off + n].view(p.shape)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g                      # autograd accumulates in place: gradients are produced inside the bucket
                self.views.append(g)
                off += n
        self.shard = torch.nn.Parameter(self.flat[rank * self.shard_len:(rank + 1) * self.shard_len], requires_grad=True)

    def collect_grads(self):
        """Gradients normally already live in the bucket; re-attach parameters whose ``.grad`` was replaced or dropped."""
        for p, g in zip(self.params, self.views):
            if p.grad is g:
                continue
            if p.grad is None:
                g.zero_()
            else:
                g.copy_(p.grad)
            p.grad = g


class TutelDistributedOptimizer:
    def __init__(self, params, group=None, average_shared=False):
        params = list(params)
        self.params = [p for p in params if not hasattr(p, '_tutel_expert')]
        self.expert_params = [p for p in params if hasattr(p, '_tutel_expert')]
        self.group = group
        self.average_shared = average_shared
        self.buckets: List[_Bucket] = []
        self.local_optim = None

    # ---- construction ------------------------------------------------------------------------------------------
    def chunk_param(self):
        """Move the dense parameters into flat buckets (one per dtype/device) and create the local shards."""
        world, rank = C.get_world_size(self.group), C.get_world_rank(self.group)
        by_kind: Dict[tuple, List[torch.nn.Parameter]] = {}
        for p in self.params:
            by_kind.setdefault((p.dtype, p.device), []).append(p)
        self.buckets = [_Bucket(ps, world, rank) for ps in by_kind.values()]

    @property
    def virt_params(self):
        """The tensors the wrapped optimizer actually updates: one flat shard per bucket."""
        return [b.shard for b in self.buckets]

    def warp_local(self, local_optim, *args, **kwargs):
        self.chunk_param()
        self.local_optim = local_optim(self.virt_params + self.expert_params, *args, **kwargs)
        return self

    wrap_local = warp_local  # correctly spelled alias of the reference's method name

    # ---- the training-step protocol ---------------------------------------------------------------------------------
    def zero_grad(self):
        for b in self.buckets:
            b.grad.zero_()
            for p, g in zip(b.params, b.views):
                p.grad = g
        for p in self.expert_params:
            if getattr(p, 'grad', None) is not None:
                p.grad.detach_()
                p.grad.zero_()

    def chunk_grad(self):
        """One reduce-scatter per bucket: this rank receives the summed gradient of its shard."""
        world = C.get_world_size(self.group)
        for b in self.buckets:
            b.collect_grads()
            g = C.simple_reduce_scatter(b.grad.view(world, b.shard_len), group=self.group).view(-1)
            if self.average_shared:
                g = g / world
            b.shard.grad = g

    def restore(self):
        """One all-gather per bucket: the updated shards land in the flat buffer every parameter is a view of."""
        for b in self.buckets:
            full = C.simple_all_gather(b.shard.data, group=self.group).view(-1)
            if full.data_ptr() != b.flat.data_ptr():
                b.flat.copy_(full)

    def step(self):
        self.chunk_grad()
        self.local_optim.step()
        self.restore()

    # ---- checkpointing of the sharded optimizer state -----------------------------------------------------------------
    def state_dict(self):
        return self.local_optim.state_dict()

    def load_state_dict(self, state):
        self.local_optim.load_state_dict(state)
