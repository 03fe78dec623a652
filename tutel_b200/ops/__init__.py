"""Operator layer: routing, sparse dispatch/combine, grouped expert GEMMs (native sm_100a kernels + CPU paths)."""
