#!/bin/bash
# Round-2 GPU call 2 (2 GPUs): multi-GPU tests (collectives, one-shot all-reduce, fused engine vs NCCL path and vs the fp32
# oracle, DP == MP), flagship bench both arms, step profile, small-message latency sweep.
OUT=gpurun_out/r2c2
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=30
timeout 600 python -m pytest tests/test_gpu_gate_route.py tests/test_gpu_kernels.py -x -q > $OUT/pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_1gpu.log
for impl in ours reference; do timeout 200 python bench/dropless_bench.py --impl $impl --megablocks_size 1 > $OUT/dropless_$impl.json 2> $OUT/dropless_$impl.err; echo "dropless $impl rc=$?"; tail -1 $OUT/dropless_$impl.json; done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
for W in coll fused oracle equiv; do
  timeout 600 $T --master-port $((29700 + RANDOM % 200)) tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error|oracle .* rel" $OUT/worker_$W.log | head -12
done
timeout 300 $T --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "ours rc=$?"; tail -c 1400 $OUT/bench_ours.json
timeout 300 $T --master-port 29912 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 900 $OUT/bench_ref.json
timeout 300 $T --master-port 29913 bench/profile_step.py --out $OUT/step_profile_2gpu.txt > $OUT/profile.log 2>&1; echo "profile rc=$?"; head -40 $OUT/step_profile_2gpu.txt
timeout 300 $T --master-port 29914 -m tutel_b200.examples.bandwidth_test --sweep --compare_nccl --loop 20 --json $OUT/a2a_sweep.json > $OUT/a2a_sweep.log 2>&1; echo "sweep rc=$?"; tail -30 $OUT/a2a_sweep.log
TUTEL_B200_SPIN_TIMEOUT_SEC=5 TUTEL_B200_FAULT='skip_push:rank=1:call=2' timeout 120 $T --master-port 29915 tests/workers/p2p_worker.py fault > $OUT/fault.log 2>&1; echo "fault rc=$? (non-zero expected)"; grep -E "FIRST_OK|timeout" $OUT/fault.log | head -5
nvidia-smi --query-gpu=index,name,memory.used --format=csv
