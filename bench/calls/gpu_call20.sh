#!/bin/bash
# Last sanity of the round: rebuilt extension (smoke), whole-step graph with the SwiGLU expert type.
OUT=gpurun_out/r2c20
mkdir -p $OUT
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 150 python bench.py --expert_type llama_ffn --steps 6 --warmup 3 > $OUT/bench_llama.json 2> $OUT/bench_llama.err; python -c "
import json
d=json.loads(open('$OUT/bench_llama.json').read().strip().splitlines()[-1]); print('llama_ffn N=1', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss0', d['first_step_loss'], 'last', d['e2e']['last_loss'], {k: v for k, v in d['config'].items() if 'graph' in k})"
