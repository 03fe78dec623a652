"""Cosine-similarity top-k gate with a learnable temperature (reference: tutel/gates/cosine_top.py:7-37).

Routing scores are cosines between a low-dimensional projection of the token and one learned prototype per expert,
sharpened by ``exp(temperature)`` (clamped so that the scale never exceeds 1 / 0.01).  Parameter names, shapes and the
order in which the initialisers draw random numbers follow the reference, so its checkpoints load and equal seeds give
equal gates.
"""
import math

import torch
import torch.nn.functional as F
from torch.nn import Linear, Parameter

_ALLOWED = ('capacity_factor', 'gate_noise')
_MAX_LOG_SCALE = math.log(1.0 / 0.01)


class CosineTopKGate(torch.nn.Module):
    def __init__(self, model_dim, num_global_experts, k=1, fp32_gate=False, proj_dim=256, init_t=0.5, **options):
        super().__init__()
        unknown = [name for name in options if name not in _ALLOWED]
        if unknown:
            raise Exception('Unrecognized argument provided to Gating module: %s' % unknown[0])
        experts = int(num_global_experts)
        self.top_k = min(experts, int(k))
        self.fp32_gate = bool(fp32_gate)
        self.clamp_max = _MAX_LOG_SCALE
        # log of the initial logit scale 1 / init_t
        self.temperature = Parameter(torch.full([1], 1.0 / init_t).log_())
        self.cosine_projector = Linear(model_dim, proj_dim)
        # two draws, like the reference (randn, then normal_(0, 0.01) over it): keeps the generator in step
        prototypes = torch.randn(proj_dim, experts)
        self.sim_matrix = Parameter(prototypes.normal_(0, 0.01))

    def logit_scale(self):
        return self.temperature.clamp(max=self.clamp_max).exp()

    def forward(self, x):
        w, b, prototypes = self.cosine_projector.weight, self.cosine_projector.bias, self.sim_matrix
        if self.fp32_gate:
            # compute in fp32 without casting the module in place (the reference calls `.float()` on the sub-module)
            x, w, b, prototypes = x.float(), w.float(), b.float(), prototypes.float()
        tokens = F.normalize(F.linear(x, w, b), dim=1)
        return (tokens @ F.normalize(prototypes, dim=0)) * self.logit_scale()


Gate = CosineTopKGate
