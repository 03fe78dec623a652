"""Built-in gates: ``top`` (linear) and ``cosine_top`` (cosine similarity router)."""
