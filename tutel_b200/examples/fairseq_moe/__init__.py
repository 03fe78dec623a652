from .integration import (MoEFeedForward, add_moe_aux_loss, convert_transformer_layers, zero_overflow_grads)  # noqa: F401
