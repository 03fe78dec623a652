"""Bias-free linear top-k gate: ``logits = x @ wg^T`` with ``wg`` of shape ``[num_global_experts, model_dim]``
(API and state-dict key ``wg.weight`` as in tutel/gates/top.py:6-25)."""
import torch
import torch.nn.functional as F


class LinearTopKGate(torch.nn.Module):
    #: per-gate options that the MoE layer consumes itself (they only have to be accepted here)
    accepted_options = frozenset({'capacity_factor', 'gate_noise'})

    def __init__(self, model_dim, num_global_experts, k=1, fp32_gate=False, **options):
        super().__init__()
        unknown = sorted(set(options) - self.accepted_options)
        if unknown:
            raise Exception('Unrecognized argument provided to Gating module: %s' % unknown[0])
        self.fp32_gate = bool(fp32_gate)
        self.top_k = min(int(k), num_global_experts)
        self.wg = torch.nn.Linear(model_dim, num_global_experts, bias=False, dtype=torch.float32 if self.fp32_gate else None)

    def forward(self, x):
        if self.fp32_gate and self.wg.weight.dtype != torch.float32:
            self.wg.float()          # the surrounding model was cast (`.half()` / `.bfloat16()`): this gate stays fp32
        weight = self.wg.weight
        return F.linear(x.to(weight.dtype), weight)


Gate = LinearTopKGate
