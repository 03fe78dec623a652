#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): fp8 inside the fused engine (tests + Mixtral-shape block bf16 vs fp8), dropless re-measure,
# GEMM shapes vs cuBLAS.
OUT=gpurun_out/r2c5
mkdir -p $OUT
export TUTEL_B200_SPIN_TIMEOUT_SEC=20
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29801 tests/workers/p2p_worker.py fp8 > $OUT/worker_fp8.log 2>&1
echo "worker fp8 rc=$? ok=$(grep -c ': OK' $OUT/worker_fp8.log) fail=$(grep -c 'FAIL' $OUT/worker_fp8.log)"; grep -E "FAIL|timeout|Error|fp8 vs bf16|Traceback" -A3 $OUT/worker_fp8.log | head -30
for v in "" "--fp8"; do
  timeout 200 $T --master-port 2981$((RANDOM % 10)) bench.py --gpus 2 --steps 10 --warmup 3 --expert_type llama_ffn $v > $OUT/mixtral_ours$v.json 2> $OUT/mixtral_ours$v.err
  echo "mixtral ours $v rc=$?"; python -c "
import json,sys
d=json.loads(open('$OUT/mixtral_ours$v.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['ms_per_step'],3),'loss',d['loss'],d['first_step_loss'])"; tail -2 $OUT/mixtral_ours$v.err
done
timeout 200 $T --master-port 29833 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 --expert_type llama_ffn --overlap 2 > $OUT/mixtral_ref.json 2> $OUT/mixtral_ref.err; python -c "
import json
d=json.loads(open('$OUT/mixtral_ref.json').read().strip().splitlines()[-1]); print('ref', round(d['ms_per_step'],3),'ms/step e2e',round(d['e2e']['ms_per_step'],3))"
for g in "" "--graph"; do timeout 100 python bench/dropless_bench.py --impl ours --megablocks_size 1 $g > $OUT/dropless_ours$g.json 2> $OUT/dropless_ours$g.err; echo "dropless ours $g rc=$?"; tail -1 $OUT/dropless_ours$g.json; done
timeout 400 python bench/kernel_check.py --inline --filter perf --out $OUT/kernel_check_perf.json > $OUT/kernel_check_perf.log 2>&1; echo "kernel_check rc=$?"; python - <<'PY'
import json
for line in open('gpurun_out/r2c5/kernel_check_perf.log'):
    if line.startswith('{'):
        r = json.loads(line); c = r['case']
        print(c.get('kind'), c.get('G'), c.get('M'), c.get('N'), c.get('K'), 'amn' if c.get('a_mn') else '', 'bmn' if c.get('b_mn') else '', 'cg', c.get('cg'),
              'ours', round(r.get('tflops_median', 0)), 'best', round(r.get('tflops_best', 0)), 'cublas', round(r.get('cublas_tflops_median', 0)), 'best', round(r.get('cublas_tflops_best', 0)), 'ok', r.get('ok'))
PY
