#!/bin/bash
# Reference arm with the final bench.py (shares the e2e loop), N=1.
OUT=gpurun_out/r2c19
mkdir -p $OUT
timeout 300 python bench.py --impl reference --steps 10 --warmup 4 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; python -c "
import json
d=json.loads(open('$OUT/bench_reference.json').read().strip().splitlines()[-1]); print('reference', d.get('impl'), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'loss0', d['first_step_loss'], 'last', d['e2e']['last_loss'], d['config'].get('cuda_graph'), d.get('unavailable'))"
tail -2 $OUT/bench_reference.err
