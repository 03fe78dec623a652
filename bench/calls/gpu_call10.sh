#!/bin/bash
# Round-2 GPU call 10 (2 GPUs): final regression - fused engine after the row-quantum change, GPU test suite, flagship N=1.
OUT=gpurun_out/r2c10
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
for W in fused deep fp8; do
  timeout 200 $T --master-port $((29700 + RANDOM % 200)) tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error" $OUT/worker_$W.log | head -6
done
timeout 100 $T --master-port 29301 -m tutel_b200.examples.helloworld_from_scratch --num_steps 6 --model_dim 512 --hidden_size 1024 --num_samples 2048 > $OUT/helloworld_from_scratch.log 2>&1; echo "from_scratch rc=$? $(grep STEP-5 $OUT/helloworld_from_scratch.log | cut -c1-80)"
CUDA_VISIBLE_DEVICES=0 timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; python -c "
import json
d=json.loads(open('$OUT/bench_ours.json').read().strip().splitlines()[-1]); print('ours N=1', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['first_step_loss'], d['gpu_launches'])"
CUDA_VISIBLE_DEVICES=0 timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
