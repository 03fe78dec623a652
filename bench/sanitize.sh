#!/bin/bash
# compute-sanitizer sweep over the native kernels (one B200; under gpurun:  gpurun --timeout 1500 -- bash bench/sanitize.sh)
# The reference has no sanitizer configuration.  Tools: memcheck (out-of-bounds / misaligned), racecheck
# (shared-memory hazards between the warp-specialised roles), synccheck (barrier misuse), initcheck (reads of
# uninitialised global memory, e.g. accumulators that a skipped tile never wrote).
# Only the small functional cases run here (the sanitizer slows kernels down 10-100x); summaries land in gpurun_out/sanitizer/.
OUT=gpurun_out/sanitizer
mkdir -p $OUT
for TOOL in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout ${TOOL_TIMEOUT:-420} compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 1 \
      python bench/kernel_check.py --no_perf --inline --out $OUT/kernel_check_$TOOL.json > $OUT/$TOOL.log 2>&1
  echo "$TOOL rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$TOOL.log | tail -1)"
done
