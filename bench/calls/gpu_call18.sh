#!/bin/bash
# bench.py e2e loop with the pipelined loss read: eager and graphed step at N=1.
OUT=gpurun_out/r2c18
mkdir -p $OUT
for g in off auto; do
  timeout 200 python bench.py --steps 20 --warmup 5 --graph $g > $OUT/bench_$g.json 2> $OUT/bench_$g.err; python -c "
import json
d=json.loads(open('$OUT/bench_$g.json').read().strip().splitlines()[-1]); print('graph=$g', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'last', d['e2e']['last_loss'], d['e2e']['d2h_bytes_per_step'], d['config'].get('cuda_graph'))"
done
tail -2 $OUT/bench_off.err
