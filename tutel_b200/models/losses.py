"""Auxiliary load-balancing losses (reference: tutel/impls/losses.py:12-42)."""
import torch
from torch.distributions.normal import Normal


def _one_hot_with_dtype(data, num_classes, dtype, hot_value=1):
    result = torch.zeros([data.size(0), num_classes], device=data.device, dtype=dtype)
    result.scatter_(1, data.unsqueeze(-1), hot_value)
    return result


def gshard_loss(scores_w_noise, top_ids):
    """GShard loss: sum_e mean-score(e) * fraction-of-first-choices(e) * E  (uses the first choice only)."""
    num_samples, num_experts = int(scores_w_noise.size(0)), int(scores_w_noise.size(1))
    first = top_ids[:, 0] if top_ids.dim() == 2 else top_ids
    # histogram without torch.bincount (which synchronises with the host to size its output)
    ce = torch.zeros([num_experts], dtype=torch.float32, device=scores_w_noise.device)
    ce.scatter_add_(0, first.reshape(-1).to(torch.int64), torch.ones([first.numel()], dtype=torch.float32, device=ce.device))
    ce = ce.to(scores_w_noise.dtype) * (num_experts / num_samples)
    me = torch.sum(scores_w_noise, dim=0)
    return torch.sum(me * ce) / num_samples


def _cv_squared(v):
    v = v.float()
    return v.var() / (v.mean() ** 2 + 1e-10)


def load_importance_loss(scores_wo_noise, topk_logits, num_global_experts, gate_noise):
    """(cv^2(importance) + cv^2(load)) / 2 with the Normal-CDF load estimate; needs ``gate_noise > 0``."""
    assert gate_noise > 0, '`gate_noise` must be > 0 for normalization in load_importance_loss().'
    device = scores_wo_noise.device
    normal = Normal(torch.tensor([0.0], device=device), torch.tensor([gate_noise / num_global_experts], device=device))
    threshold = topk_logits[:, -1].reshape(-1, 1).float()
    load = normal.cdf(scores_wo_noise.float() - threshold).sum(0)
    importance = scores_wo_noise.float().sum(0)
    return (_cv_squared(importance) + _cv_squared(load)) / 2.0
