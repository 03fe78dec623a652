"""ZeRO-1/2 style optimizer wrapper (mirrors ``TutelDistributedOptimizer``, tutel/net.py:15-58).

Dense (non-expert) parameters: optimizer state lives on a flat 1/W shard per rank - gradients are reduce-scattered,
the local optimizer steps its shard, updated parameters are all-gathered.  Expert parameters (tagged
``_tutel_expert`` by the MoE layer) are stepped locally.
"""
from . import communicate as C


class TutelDistributedOptimizer:
    def __init__(self, params, group=None, average_shared=False):
        params = list(params)
        self.params = [p for p in params if not hasattr(p, '_tutel_expert')]
        self.expert_params = [p for p in params if hasattr(p, '_tutel_expert')]
        self.shapes = [p.shape for p in self.params]
        self.group = group
        self.average_shared = average_shared
        self.virt_params = []
        self.local_optim = None

    def chunk_param(self):
        self.virt_params = [C.zero_scatter(p.data, C.simple_split, group=self.group)[0] for p in self.params]

    def chunk_grad(self):
        world = C.get_world_size(self.group)
        for shard, p in zip(self.virt_params, self.params):
            if getattr(p, 'grad', None) is None:
                continue
            grad = p.grad.reshape(-1)
            if self.average_shared:
                grad = grad / world
            shard.grad, _ = C.zero_scatter(grad, C.simple_reduce_scatter, group=self.group)

    def restore(self):
        for shard, p, shape in zip(self.virt_params, self.params, self.shapes):
            full = C.simple_all_gather(shard.data, group=self.group).view(-1)
            p.data = full[: shape.numel()].view(shape)

    def warp_local(self, local_optim, *args, **kwargs):
        self.chunk_param()
        self.local_optim = local_optim(self.virt_params + self.expert_params, *args, **kwargs)
        return self

    wrap_local = warp_local  # correctly spelled alias

    def zero_grad(self):
        for p in self.params + self.expert_params:
            if getattr(p, 'grad', None) is not None:
                p.grad.detach_()
                p.grad.zero_()

    def step(self):
        self.chunk_grad()
        self.local_optim.step()
        self.restore()
