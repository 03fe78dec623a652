"""Public MoE API (mirrors tutel/moe.py:6-12)."""
# Low-level ops
from .ops.gating import fast_cumsum_sub_one
from .ops.dispatch import fast_dispatcher, fast_encode, fast_decode
from .ops.routing import extract_critical

top_k_routing = extract_critical

# High-level op
from .models.moe_layer import moe_layer

__all__ = ['moe_layer', 'top_k_routing', 'extract_critical', 'fast_encode', 'fast_decode', 'fast_dispatcher',
           'fast_cumsum_sub_one']
