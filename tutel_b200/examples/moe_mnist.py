#!/usr/bin/env python3
"""MNIST with an MoE classifier head (reference: tutel/examples/moe_mnist.py).
    python -m tutel_b200.examples.moe_mnist --epochs 2        (add --no-moe for the dense baseline)"""
from tutel_b200.examples._vision import run

if __name__ == '__main__':
    run('mnist')
