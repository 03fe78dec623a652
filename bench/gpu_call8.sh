#!/bin/bash
# Round-2 8-GPU evidence run (under `gpurun --gpus 8`): worker suite, flagship bench both arms, Mixtral-shape block
# (bf16 fused / fp8 fused / reference bf16 overlap 2), 16-layer stack in a 4 GiB arena, all-to-all / all-reduce sweep.
N=${1:-8}
OUT=gpurun_out/r2_scale$N
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=30
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
P=29600
run() { name=$1; shift; P=$((P+1)); timeout 200 $T --master-port $P "$@" > $OUT/$name.json 2> $OUT/$name.err; tail -c 3000 $OUT/$name.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'], 3), 'tok/s', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'], 3) if d.get('e2e') else None, 'first loss', d.get('first_step_loss'), d.get('clocks', {}).get('reasons'))
except Exception as ex:
    print('$name FAILED', ex)
"; grep -E "timeout|Error" $OUT/$name.err | head -3; }
for W in coll fused oracle deep fp8; do
  P=$((P+1)); timeout 300 $T --master-port $P tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error" $OUT/worker_$W.log | head -6
done
run bench_ours bench.py --gpus $N --steps 20 --warmup 5
run bench_reference bench.py --impl reference --gpus $N --steps 20 --warmup 5
run mixtral_ours_bf16 bench.py --gpus $N --steps 10 --warmup 3 --expert_type llama_ffn
run mixtral_ours_fp8 bench.py --gpus $N --steps 10 --warmup 3 --expert_type llama_ffn --fp8
run mixtral_reference_bf16_d2 bench.py --impl reference --gpus $N --steps 10 --warmup 3 --expert_type llama_ffn --overlap 2
P=$((P+1)); TUTEL_B200_HEAP_MB=4096 TUTEL_B200_STAGE_MB=1024 timeout 200 $T --master-port $P bench/deep_stack.py --layers 16 --steps 3 > $OUT/deep_stack16.json 2> $OUT/deep_stack16.err; echo "deep rc=$?"; tail -1 $OUT/deep_stack16.json
P=$((P+1)); timeout 200 $T --master-port $P -m tutel_b200.examples.bandwidth_test --sweep --compare_nccl --loop 20 --json $OUT/a2a_sweep.json > $OUT/a2a_sweep.log 2>&1; echo "sweep rc=$?"; grep -E "all_to_all|all_reduce" $OUT/a2a_sweep.log | head -60
