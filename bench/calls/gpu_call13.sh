#!/bin/bash
# MX GEMM on CTA pairs (cta_group::2): exactness cases, then throughput next to the 1-CTA kernel; MX pytest file.
mkdir -p gpurun_out/mx3
timeout 600 python bench/mx_check.py --out gpurun_out/mx3 --only exact_cg2,exact_cg2_small_grid 2>&1 | tee gpurun_out/mx3/run.log | tail -40
timeout 250 python -m pytest tests/test_gpu_mx.py -x -q 2>&1 | tail -15 | tee gpurun_out/mx3/pytest.log
