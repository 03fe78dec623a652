#!/bin/bash
# Round-2 GPU call 6 (1 GPU): per-kernel profile of the fp8 SwiGLU block, ncu of the hot single-GPU kernels, sanitizers.
OUT=gpurun_out/r2c6
mkdir -p $OUT
timeout 200 python bench/profile_step.py --expert_type llama_ffn --fp8 --steps 3 --out $OUT/step_profile_llama_fp8.txt > $OUT/profile_fp8.log 2>&1; echo "profile fp8 rc=$?"; head -24 $OUT/step_profile_llama_fp8.txt | cut -c1-170
timeout 200 python bench/profile_step.py --expert_type llama_ffn --steps 3 --out $OUT/step_profile_llama_bf16.txt > $OUT/profile_bf16.log 2>&1; echo "profile bf16 rc=$?"; head -12 $OUT/step_profile_llama_bf16.txt | cut -c1-170
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/kernels python bench/ncu_targets.py > $OUT/ncu_kernels.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu_kernels.log
TOOLS="memcheck racecheck" TOOL_TIMEOUT=300 bash bench/sanitize.sh; cp -r gpurun_out/sanitizer $OUT/ 2>/dev/null
ls -la $OUT | head -30
