#!/bin/bash
# MX path, second pass: persistent kernel with register hand-off epilogue, bias / ReLU / ReLU-backward epilogues,
# transposing quantiser, autograd FFN, layer integration; flagship step with MX experts next to row-scaled fp8.
mkdir -p gpurun_out/mx2
timeout 700 python bench/mx_check.py --out gpurun_out/mx2 2>&1 | tee gpurun_out/mx2/run.log | tail -45
timeout 250 python -m pytest tests/test_gpu_mx.py -x -q 2>&1 | tail -12 | tee gpurun_out/mx2/pytest.log
for mode in mx row; do
  timeout 200 python bench.py --fp8 --fp8_mode $mode --steps 10 --warmup 4 2>gpurun_out/mx2/bench_$mode.err | tail -1 | tee gpurun_out/mx2/bench_$mode.json | cut -c1-400
done
