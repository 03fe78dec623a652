#!/bin/bash
# Round-2 8-GPU evidence run (under `gpurun --gpus 8`): flagship bench both arms (N=8, N=4), Mixtral-shape block
# (bf16 fused / fp8 fused / reference bf16 overlap 2), 16-layer stack in a 4 GiB arena, worker suite incl. sub-groups and the
# hierarchical all-to-all, all-to-all / all-reduce sweep.
OUT=gpurun_out/r2_scale8
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=30
P=29600
TN() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1"; }
run() { n=$1; name=$2; shift 2; P=$((P+1)); timeout 150 $(TN $n) --master-port $P "$@" > $OUT/$name.json 2> $OUT/$name.err; tail -c 3000 $OUT/$name.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$name', 'ms/step', round(d['ms_per_step'], 3), 'tok/s', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'], 3) if d.get('e2e') else None, 'first loss', d.get('first_step_loss'), 'loss', d.get('loss'), d.get('clocks', {}).get('reasons'))
except Exception as ex:
    print('$name FAILED', ex)
"; grep -E "timeout|Error" $OUT/$name.err | head -3; }
run 8 bench_ours bench.py --gpus 8 --steps 20 --warmup 5
run 8 bench_reference bench.py --impl reference --gpus 8 --steps 20 --warmup 5
run 8 mixtral_ours_bf16 bench.py --gpus 8 --steps 10 --warmup 3 --expert_type llama_ffn
run 8 mixtral_ours_fp8 bench.py --gpus 8 --steps 10 --warmup 3 --expert_type llama_ffn --fp8
run 8 mixtral_reference_bf16_d2 bench.py --impl reference --gpus 8 --steps 10 --warmup 3 --expert_type llama_ffn --overlap 2
run 8 flagship_ours_fp8 bench.py --gpus 8 --steps 10 --warmup 3 --fp8
run 4 bench_ours_n4 bench.py --gpus 4 --steps 20 --warmup 5
run 4 bench_reference_n4 bench.py --impl reference --gpus 4 --steps 20 --warmup 5
P=$((P+1)); TUTEL_B200_HEAP_MB=4096 TUTEL_B200_STAGE_MB=1024 timeout 150 $(TN 8) --master-port $P bench/deep_stack.py --layers 16 --steps 3 > $OUT/deep_stack16.json 2> $OUT/deep_stack16.err; echo "deep rc=$?"; tail -1 $OUT/deep_stack16.json; grep -E "Error|timeout" $OUT/deep_stack16.err | head -3
for W in coll fused oracle deep fp8 sub; do
  P=$((P+1)); TUTEL_B200_TEST_FLAGSHIP=0 timeout 200 $(TN 8) --master-port $P tests/workers/p2p_worker.py $W > $OUT/worker_$W.log 2>&1
  echo "worker $W rc=$? ok=$(grep -c ': OK' $OUT/worker_$W.log) fail=$(grep -c 'FAIL' $OUT/worker_$W.log)"; grep -E "FAIL|timeout|Error" $OUT/worker_$W.log | head -6
done
P=$((P+1)); LOCAL_SIZE=2 timeout 150 $(TN 8) --master-port $P tests/workers/p2p_worker.py 2dh > $OUT/worker_2dh.log 2>&1; echo "worker 2dh rc=$? ok=$(grep -c ': OK' $OUT/worker_2dh.log) fail=$(grep -c 'FAIL' $OUT/worker_2dh.log)"; grep -E "FAIL|timeout|Error" $OUT/worker_2dh.log | head -6
P=$((P+1)); timeout 150 $(TN 8) --master-port $P -m tutel_b200.examples.bandwidth_test --sweep --compare_nccl --loop 10 --json $OUT/a2a_sweep.json > $OUT/a2a_sweep.log 2>&1; echo "sweep rc=$?"; grep -E "^(all_to_all|all_reduce|nccl_all)" $OUT/a2a_sweep.log | head -50
