#!/bin/bash
# Kernel timelines (CUPTI through torch.profiler) of the graphed step at 8 views and at 1 view per step (= a rank's ray count
# at 8 GPUs): where the step's time goes BETWEEN the kernels.   usage: bash profiles/timeline_session.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
COMMON="--no-cpu-baseline --no-ref-cuda --no-shading --profile-steps 2"
timeout 600 python bench.py --steps 30 --warmup 8 $COMMON --timeline $O/timeline_${TAG}_v8.json > $O/bench_${TAG}_tl_v8.json 2> $O/bench_${TAG}_tl_v8.err
echo "v8 exit $?"; cut -c1-200 $O/bench_${TAG}_tl_v8.json
timeout 600 python bench.py --views 1 --steps 100 --warmup 20 $COMMON --timeline $O/timeline_${TAG}_v1.json > $O/bench_${TAG}_tl_v1.json 2> $O/bench_${TAG}_tl_v1.err
echo "v1 exit $?"; cut -c1-200 $O/bench_${TAG}_tl_v1.json
timeout 600 python bench.py --views 1 --pipeline --steps 100 --warmup 20 $COMMON --timeline $O/timeline_${TAG}_v1p.json > $O/bench_${TAG}_tl_v1p.json 2> $O/bench_${TAG}_tl_v1p.err
echo "v1 pipelined exit $?"; cut -c1-200 $O/bench_${TAG}_tl_v1p.json
