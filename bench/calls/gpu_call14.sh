#!/bin/bash
# MX evidence: ncu --set full of the MX GEMM (CTA pairs, single CTA), the row-scaled fp8 GEMM and the MX quantisers at the
# flagship expert shape; full mx_check (exactness + throughput table + expert-FFN forward/backward); flagship step with MX experts.
mkdir -p gpurun_out/mx4
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/mx4/mx_ncu python bench/prof_mx.py > gpurun_out/mx4/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/mx4/ncu.log
timeout 100 python bench/ncu_summary.py gpurun_out/mx4/mx_ncu.ncu-rep > gpurun_out/mx4/mx_ncu_summary.txt 2>&1; head -50 gpurun_out/mx4/mx_ncu_summary.txt
timeout 600 python bench/mx_check.py --out gpurun_out/mx4 2>&1 | tee gpurun_out/mx4/run.log | grep -E "^[a-z_0-9]+ \{|shape|ffn_shape" | cut -c1-900
timeout 200 python bench.py --fp8 --fp8_mode mx --steps 10 --warmup 4 2>gpurun_out/mx4/bench_mx.err | tail -1 | tee gpurun_out/mx4/bench_mx.json | cut -c1-300
