#!/bin/bash
# Whole-step CUDA graph (GraphedTrainStep) at N=1: GPU test, flagship bench with and without the graph, fp8 modes under the graph.
mkdir -p gpurun_out/graph
timeout 200 python -m pytest tests/test_gpu_gate_route.py -x -q -k "graphed" 2>&1 | tail -5 | tee gpurun_out/graph/pytest.log
for g in auto off; do
  timeout 250 python bench.py --steps 20 --warmup 5 --graph $g 2>gpurun_out/graph/bench_$g.err | tail -1 > gpurun_out/graph/bench_$g.json
  python - <<PY
import json
d = json.load(open('gpurun_out/graph/bench_$g.json'))
print('$g', 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches'], 'loss0', d['first_step_loss'], 'last', d['e2e']['last_loss'],
      {k: v for k, v in d['config'].items() if 'graph' in k}, d.get('clocks'))
PY
done
for m in row mx; do
  timeout 200 python bench.py --steps 10 --warmup 4 --fp8 --fp8_mode $m 2>gpurun_out/graph/bench_fp8_$m.err | tail -1 > gpurun_out/graph/bench_fp8_$m.json
  python - <<PY
import json
d = json.load(open('gpurun_out/graph/bench_fp8_$m.json'))
print('fp8 $m', 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), {k: v for k, v in d['config'].items() if 'graph' in k})
PY
done
tail -3 gpurun_out/graph/bench_auto.err
