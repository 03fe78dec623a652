"""Auto-SPMD search (``tutel.parted``): describe a computation as einsum-like ops over named tensors, let the solver
pick a sharding state per tensor and a collective pattern per op by *measuring* generated programs."""
