#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): transposing quantisation, fp8 engine re-check, Mixtral-shape block bf16 vs fp8.
OUT=gpurun_out/r2c7
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
timeout 300 python -m pytest tests/test_gpu_gate_route.py tests/test_gpu_kernels.py -x -q > $OUT/pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_1gpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29801 tests/workers/p2p_worker.py fp8 > $OUT/worker_fp8.log 2>&1
echo "worker fp8 rc=$? ok=$(grep -c ': OK' $OUT/worker_fp8.log) fail=$(grep -c 'FAIL' $OUT/worker_fp8.log)"; grep -E "FAIL|timeout|Error|Traceback" -A3 $OUT/worker_fp8.log | head -20
for v in "" "--fp8"; do
  timeout 200 $T --master-port 2981$((RANDOM % 10)) bench.py --gpus 2 --steps 10 --warmup 3 --expert_type llama_ffn $v > $OUT/mixtral_ours$v.json 2> $OUT/mixtral_ours$v.err
  echo "mixtral ours $v rc=$?"; python -c "
import json,sys
d=json.loads(open('$OUT/mixtral_ours$v.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['ms_per_step'],3),'loss',d['loss'],d['first_step_loss'])"; tail -2 $OUT/mixtral_ours$v.err
done
timeout 200 $T --master-port 29855 bench.py --gpus 2 --steps 10 --warmup 3 --fp8 > $OUT/flagship_fp8.json 2> $OUT/flagship_fp8.err; python -c "
import json
d=json.loads(open('$OUT/flagship_fp8.json').read().strip().splitlines()[-1]); print('flagship fp8', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['ms_per_step'],3), d['loss'], d['first_step_loss'])"
timeout 200 python bench/profile_step.py --expert_type llama_ffn --fp8 --steps 3 --out $OUT/step_profile_llama_fp8.txt > $OUT/profile_fp8.log 2>&1; echo "profile fp8 rc=$?"; head -12 $OUT/step_profile_llama_fp8.txt | cut -c1-170
