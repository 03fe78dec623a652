#!/bin/bash
# Final regression of round 2 (1 GPU): full GPU test suite, smoke, flagship bench (ours), memcheck of the MX kernels.
OUT=gpurun_out/r2c17
mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 250 python bench.py --steps 20 --warmup 5 > $OUT/bench_ours.json 2> $OUT/bench_ours.err; python -c "
import json
d=json.loads(open('$OUT/bench_ours.json').read().strip().splitlines()[-1]); print('ours N=1', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['first_step_loss'], d['gpu_launches'], d['config'].get('cuda_graph'), d.get('clocks'))"
for grp in "exact --quick 1" "exact_cg2 --small_grid 1"; do
  timeout 200 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 1 python bench/mx_check.py --group $grp > $OUT/memcheck_mx_$(echo $grp | cut -d' ' -f1).log 2>&1
  echo "memcheck $grp rc=$? $(grep -E 'ERROR SUMMARY' $OUT/memcheck_mx_$(echo $grp | cut -d' ' -f1).log | tail -1) $(grep -c '"mismatch": 0' $OUT/memcheck_mx_$(echo $grp | cut -d' ' -f1).log) exact"
done
