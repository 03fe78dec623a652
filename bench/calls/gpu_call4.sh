#!/bin/bash
# Round-2 GPU call 4 (1 GPU): full GPU test suite, dropless eager/graph vs reference, flagship bench (no per-forward host sync).
OUT=gpurun_out/r2c4
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
for g in "" "--graph"; do timeout 100 python bench/dropless_bench.py --impl ours --megablocks_size 1 $g > $OUT/dropless_ours$g.json 2> $OUT/dropless_ours$g.err; echo "dropless ours $g rc=$?"; tail -1 $OUT/dropless_ours$g.json; tail -2 $OUT/dropless_ours$g.err; done
timeout 100 python bench/dropless_bench.py --impl reference --megablocks_size 1 > $OUT/dropless_reference.json 2> $OUT/dropless_reference.err; tail -1 $OUT/dropless_reference.json
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "ours rc=$?"; tail -c 1500 $OUT/bench_ours.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 700 $OUT/bench_ref.json
timeout 200 python bench/profile_step.py --out $OUT/step_profile_1gpu.txt > $OUT/profile.log 2>&1; echo "profile rc=$?"; head -12 $OUT/step_profile_1gpu.txt | cut -c1-150
