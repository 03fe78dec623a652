#!/usr/bin/env python3
"""CIFAR-10 with an MoE classifier head (reference: tutel/examples/moe_cifar10.py).
    python -m tutel_b200.examples.moe_cifar10 --epochs 2"""
from tutel_b200.examples._vision import run

if __name__ == '__main__':
    run('cifar10')
