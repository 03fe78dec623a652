"""Parallel runtime: process groups, collectives (NCCL/Gloo and NVLink P2P), overlap scheduling, ZeRO optimizer."""
