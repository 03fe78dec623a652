#!/bin/bash
# 2-GPU regression after the MX / CUDA-graph work: fused engine checks, fp8 (row + mx) worker, flagship bench at N=2 (ours).
OUT=gpurun_out/r2c16
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
for W in fused fp8; do
  timeout 200 $T --master-port $((29700 + RANDOM % 200)) tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error|mx vs bf16" $OUT/worker_$W.log | head -6
done
timeout 200 $T --master-port 29311 bench.py --gpus 2 --steps 10 --warmup 4 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; python -c "
import json
d=json.loads(open('$OUT/bench_n2.json').read().strip().splitlines()[-1]); print('ours N=2', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['first_step_loss'], d['gpu_launches'], d['config'].get('cuda_graph'))"
