#!/bin/bash
# One development iteration on the GPU: the GPU tests, then short bench runs (8 views, 1 view) with kernel timelines.
#   usage: bash profiles/iter_session.sh <tag> [tests|notests] [extra pytest args]
TAG=${1:-it}
DO_TESTS=${2:-tests}
O=gpurun_out
mkdir -p $O
if [ "$DO_TESTS" = tests ]; then
  ( time timeout 900 python -m pytest tests -m gpu -q -x ${3:-} ) > $O/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
  tail -15 $O/pytest_gpu_$TAG.log
fi
COMMON="--no-cpu-baseline --no-ref-cuda --no-shading --profile-steps 2"
timeout 600 python bench.py --steps 48 --warmup 8 $COMMON --timeline $O/timeline_${TAG}_v8.json > $O/bench_${TAG}_v8.json 2> $O/bench_${TAG}_v8.err
echo "v8 exit $?"; cut -c1-200 $O/bench_${TAG}_v8.json; tail -3 $O/bench_${TAG}_v8.err
timeout 600 python bench.py --views 1 --steps 128 --warmup 20 $COMMON --timeline $O/timeline_${TAG}_v1.json > $O/bench_${TAG}_v1.json 2> $O/bench_${TAG}_v1.err
echo "v1 exit $?"; cut -c1-200 $O/bench_${TAG}_v1.json; tail -3 $O/bench_${TAG}_v1.err
