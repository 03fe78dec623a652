"""PyTorch code-generation backend of the auto-SPMD tool."""
