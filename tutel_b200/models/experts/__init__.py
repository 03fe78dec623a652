"""Built-in experts: ``ffn`` (2-layer, biases, hidden-dim sharding) and ``llama_ffn`` (SwiGLU, flat-sharded)."""
