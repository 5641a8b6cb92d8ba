#!/bin/bash
# Profiling recipe (B200_PROFILING.md), run under gpurun on ONE GPU, in two calls (a call may bring back at most 64 MiB):
#   profiles/run_ncu.sh <tag> default     launch list + k_trace_both + k_shade_queues of the default schedule
#   profiles/run_ncu.sh <tag> separate    k_trace_closest + k_shadow_accumulate as separate kernels (--overlap 0), the
#                                         configuration bench.py's roofline region times
# then, back in the build container:  python profiles/make_summaries.py <tag>
set -x
TAG=${1:-r01}
PART=${2:-default}
mkdir -p gpurun_out
if [ "$PART" = default ]; then
# launch list of the default bench command with per-launch device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench_${TAG}.log 2>&1
# the counter frame (COUNT variants) does not use k_trace_both: the first 8 launches are bounces 1..8 of the first warm-up frame
ncu --set full --clock-control none --import-source on -k regex:k_trace_both -s 0 -c 8 -f -o gpurun_out/prof_both_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_both_bench_${TAG}.log 2>&1
# k_shade_queues has no COUNT variant: the counter frame launches it too (9 launches) -> skip them
ncu --set full --clock-control none --import-source on -k regex:k_shade_queues -s 9 -c 9 -f -o gpurun_out/prof_shade_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shade_bench_${TAG}.log 2>&1
else
# skip the 9 instrumented (COUNT) launches of the counter frame, capture all 9 bounces of the first warm-up frame
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 9 -c 9 -f -o gpurun_out/prof_trace_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --overlap 0 > gpurun_out/ncu_trace_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shadow_accumulate -s 9 -c 9 -f -o gpurun_out/prof_shadow_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --overlap 0 > gpurun_out/ncu_shadow_bench_${TAG}.log 2>&1
fi
ls -la gpurun_out
