"""fairseq integration without patching fairseq.

The reference ships ``fairseq_patch.diff`` (tutel/examples/fairseq_moe/fairseq_patch.diff:24-128), which edits fairseq's
``TransformerDecoderLayerBase`` so that every ``MOE``-th layer builds a ``tutel.moe.moe_layer`` instead of ``fc1/fc2``,
records ``(num_tokens, l_aux)`` per layer in ``tutel.system.cache()`` and lets ``FairseqTask.train_step`` add
``L_AUX_WT * sum(samples * l_aux)`` to the loss; ``NO_OVERFLOW=1`` zeroes inf gradients in the fp16 optimizer.

Here the same behaviour is obtained by converting an already-built model in place - no source patch, so it works
with any fairseq version whose transformer layers expose ``fc1, fc2, activation_fn, activation_dropout_module`` (and
optionally ``ffn_layernorm``):

    model = task.build_model(cfg.model)
    convert_transformer_layers(model)                      # honours MOE=<freq> like the reference patch
    ...
    loss = add_moe_aux_loss(loss)                          # honours L_AUX_WT; call once per train step
    zero_overflow_grads(optimizer_params)                  # honours NO_OVERFLOW (fp16)

The conversion keeps fairseq's own ``forward`` untouched: ``fc1`` becomes the whole MoE feed-forward, the inline
activation / dropout / layer-norm / ``fc2`` that follow it in fairseq's forward become identities, and their original
modules run inside the experts' activation exactly where the reference patch puts them.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, List, Optional

import torch

from ... import moe as tutel_moe
from ... import system


class _Identity(torch.nn.Module):
    def forward(self, x, *args, **kwargs):
        return x


class MoEFeedForward(torch.nn.Module):
    """Stands in for ``fc1`` of a fairseq transformer layer and computes the complete MoE feed-forward."""

    def __init__(self, embed_dim: int, ffn_dim: int, inner: Callable[[torch.Tensor], torch.Tensor],
                 num_experts_per_device: int = 1, top_k: int = 2, group=None):
        super().__init__()
        self.moe_ffn = tutel_moe.moe_layer(
            gate_type={'type': 'top', 'k': top_k, 'capacity_factor': 0.0, 'fp32_gate': True, 'gate_noise': 1.0},
            model_dim=embed_dim,
            experts={'type': 'ffn', 'num_experts_per_device': num_experts_per_device, 'hidden_size_per_expert': ffn_dim,
                     'activation_fn': inner},
            # same marker as the reference patch: fairseq's legacy_ddp skips parameters tagged `expert`
            scan_expert_func=lambda name, param: setattr(param, 'expert', True), group=group)

    def forward(self, x):
        y = self.moe_ffn(x)
        if y.l_aux is not None and getattr(y.l_aux, 'requires_grad', False):
            system.cache().set(id(self.moe_ffn), (x.numel() // x.size(-1), y.l_aux))
        return y


def _ffn_layers(model: torch.nn.Module) -> List[torch.nn.Module]:
    need = ('fc1', 'fc2', 'activation_fn')
    return [m for m in model.modules() if all(hasattr(m, a) for a in need) and isinstance(getattr(m, 'fc1'), torch.nn.Module)]


def convert_transformer_layers(model: torch.nn.Module, moe_freq: Optional[int] = None, num_experts_per_device: int = 1,
                               top_k: int = 2, group=None) -> int:
    """Replace the FFN of every ``moe_freq``-th transformer layer (1-based, like the patch: ``(index + 1) % freq == 0``)
    by a ``tutel_b200`` MoE layer.  ``moe_freq`` defaults to the ``MOE`` environment variable; 0 converts nothing.
    Returns the number of converted layers."""
    freq = int(os.environ.get('MOE', 0)) if moe_freq is None else int(moe_freq)
    if freq <= 0:
        return 0
    converted = 0
    for index, layer in enumerate(_ffn_layers(model)):
        if (index + 1) % freq != 0:
            continue
        assert float(getattr(layer, 'quant_noise', 0) or 0) == 0, 'Unhandled quant_noise > 0.0 for MoE layer.'
        fc1, fc2 = layer.fc1, layer.fc2
        embed_dim = getattr(layer, 'embed_dim', None) or fc1.in_features
        ffn_dim = fc1.out_features
        act = layer.activation_fn
        drop = getattr(layer, 'activation_dropout_module', None) or _Identity()
        norm = getattr(layer, 'ffn_layernorm', None)

        def inner(h, act=act, drop=drop, norm=norm):
            h = drop(act(h))
            return h if norm is None else norm(h)

        ref = fc1.weight
        stage = MoEFeedForward(embed_dim, ffn_dim, inner, num_experts_per_device, top_k, group).to(ref.device).to(ref.dtype)
        if norm is not None:
            stage.ffn_layernorm = norm        # keep its parameters registered (and trained) under the MoE stage
            layer.ffn_layernorm = None
        layer.fc1 = stage
        layer.fc2 = _Identity()
        layer.activation_fn = lambda t: t
        if hasattr(layer, 'activation_dropout_module'):
            layer.activation_dropout_module = _Identity()
        converted += 1
    return converted


def add_moe_aux_loss(loss: torch.Tensor, l_aux_wt: Optional[float] = None) -> torch.Tensor:
    """``loss + L_AUX_WT * sum_layers(num_tokens * l_aux)`` from the per-layer records of this step; clears the records."""
    wt = float(os.environ.get('L_AUX_WT', 0.0)) if l_aux_wt is None else float(l_aux_wt)
    cache = system.cache()
    if wt:
        total = None
        for samples, l_aux in cache.get():
            term = l_aux * (wt * samples)
            total = term if total is None else total + term
        if total is not None:
            loss = loss + total.to(loss.dtype)
    cache.reset()
    return loss


def zero_overflow_grads(params: Iterable[torch.nn.Parameter], enabled: Optional[bool] = None) -> None:
    """fp16 training helper (``NO_OVERFLOW=1`` in the reference patch): replace infinite gradient entries by zero."""
    on = int(os.environ.get('NO_OVERFLOW', 0)) > 0 if enabled is None else enabled
    if not on:
        return
    for p in params:
        if p.grad is not None:
            p.grad.masked_fill_(torch.isinf(p.grad), 0)
