#!/bin/bash
# Round-2 GPU call 1 (1 GPU): GPU tests, flagship bench both arms, per-kernel step profile, ncu of the main GEMM.
OUT=gpurun_out/r2c1
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "ours rc=$?"; tail -c 1500 $OUT/bench_ours.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 1200 $OUT/bench_ref.json
timeout 300 python bench/profile_step.py --out $OUT/step_profile_1gpu.txt > $OUT/profile.log 2>&1; echo "profile rc=$?"; head -30 $OUT/step_profile_1gpu.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_sm100 --profile-from-start off -c 1 -o $OUT/gemm_cg2 python bench/prof_gemm.py cg2 > $OUT/ncu_gemm.log 2>&1; echo "ncu rc=$?"
ls -la $OUT
