"""Cosine-similarity top-k gate with learnable temperature (reference: tutel/gates/cosine_top.py:7-37)."""
import math

import torch
import torch.nn.functional as F

_ALLOWED = ('capacity_factor', 'gate_noise')


class CosineTopKGate(torch.nn.Module):
    def __init__(self, model_dim, num_global_experts, k=1, fp32_gate=False, proj_dim=256, init_t=0.5, **options):
        super().__init__()
        for opt in options:
            if opt not in _ALLOWED:
                raise Exception('Unrecognized argument provided to Gating module: %s' % opt)
        self.top_k = min(num_global_experts, int(k))
        self.fp32_gate = fp32_gate
        self.temperature = torch.nn.Parameter(torch.log(torch.full([1], 1.0 / init_t)), requires_grad=True)
        self.cosine_projector = torch.nn.Linear(model_dim, proj_dim)
        self.sim_matrix = torch.nn.Parameter(torch.randn(size=(proj_dim, num_global_experts)), requires_grad=True)
        self.clamp_max = math.log(1.0 / 0.01)
        torch.nn.init.normal_(self.sim_matrix, 0, 0.01)

    def forward(self, x):
        projector, sim = self.cosine_projector, self.sim_matrix
        if self.fp32_gate:
            x, projector, sim = x.float(), projector.float(), sim.float()
        logits = torch.matmul(F.normalize(projector(x), dim=1), F.normalize(sim, dim=0))
        return logits * torch.clamp(self.temperature, max=self.clamp_max).exp()


Gate = CosineTopKGate
