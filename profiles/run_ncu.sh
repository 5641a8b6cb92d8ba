#!/bin/bash
# Profiling recipe (B200_PROFILING.md), run under gpurun on ONE GPU:
#   profiles/run_ncu.sh <tag>
# 1. launch list with per-launch device time (cold-cache, serialised: compare SHARES)
# 2. full capture of the dominant kernels (fused extend+shade bounce 0/1, shadow+accumulate bounce 0)
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench_${TAG}.log 2>&1
# skip the 9 instrumented (COUNT) launches of the counter frame, capture all 9 bounces of the first warm-up frame
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 9 -c 9 -f -o gpurun_out/prof_trace_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_trace_bench_${TAG}.log 2>&1
# k_shade_queues has no COUNT variant: the counter frame launches it too (9 launches) -> skip them as well
ncu --set full --clock-control none --import-source on -k regex:k_shade_queues -s 9 -c 9 -f -o gpurun_out/prof_shade_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shade_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shadow_accumulate -s 9 -c 9 -f -o gpurun_out/prof_shadow_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shadow_bench_${TAG}.log 2>&1
ls -la gpurun_out
