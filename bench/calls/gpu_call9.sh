#!/bin/bash
# Round-2 GPU call 9 (2 GPUs): regression after the last changes (all-reduce thresholds, table caching, GC-free timing) and
# every example program for a few steps.
OUT=gpurun_out/r2c9
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
for W in coll fused; do
  timeout 200 $T --master-port $((29700 + RANDOM % 200)) tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error" $OUT/worker_$W.log | head -6
done
timeout 150 $T --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; python -c "
import json
d=json.loads(open('$OUT/bench_ours.json').read().strip().splitlines()[-1]); print('ours N=2', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['first_step_loss'])"
bash bench/examples_smoke.sh 2
cp -r gpurun_out/examples_2 $OUT/ 2>/dev/null
