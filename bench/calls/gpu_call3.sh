#!/bin/bash
# Round-2 GPU call 3 (2 GPUs, short): new fused engine (ring / leases / all modes / deep stack), oracle at flagship shape,
# bench both arms, dropless with CUDA graph.
OUT=gpurun_out/r2c3
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
timeout 300 python -m pytest tests/test_gpu_gate_route.py tests/test_gpu_kernels.py -x -q > $OUT/pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_1gpu.log
for g in "" "--graph"; do timeout 100 python bench/dropless_bench.py --impl ours --megablocks_size 1 $g > $OUT/dropless_ours$g.json 2> $OUT/dropless_ours$g.err; echo "dropless ours $g rc=$?"; tail -1 $OUT/dropless_ours$g.json; done
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
for W in fused deep oracle equiv; do
  timeout 240 $T --master-port $((29700 + RANDOM % 200)) tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error|rel err" $OUT/worker_$W.log | head -12
done
timeout 200 $T --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "ours rc=$?"; tail -c 700 $OUT/bench_ours.json; tail -3 $OUT/bench_ours.err
TUTEL_B200_SPIN_TIMEOUT_SEC=5 TUTEL_B200_FAULT='skip_push:rank=1:call=2' timeout 90 $T --master-port 29915 tests/workers/p2p_worker.py fault > $OUT/fault.log 2>&1; echo "fault rc=$? (non-zero expected)"; grep -E "FIRST_OK|timeout" $OUT/fault.log | head -3
