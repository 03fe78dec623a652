#!/bin/bash
# MX block-scaled GEMM bring-up: crafted exactness cases (both scale-address variants), random data, throughput.
mkdir -p gpurun_out/mx
timeout 700 python bench/mx_check.py --out gpurun_out/mx 2>&1 | tee gpurun_out/mx/run.log | tail -60
timeout 200 python -m pytest tests/test_gpu_mx.py -x -q 2>&1 | tail -8 | tee gpurun_out/mx/pytest.log
