#!/bin/bash
# Last check of the final build (1 GPU): GPU tests, smoke, the default bench line, the reference arm, infer and encoder lines.
TAG=${1:-r2_final2}
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
tail -4 $O/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $O/smoke_$TAG.log
tail -2 $O/smoke_$TAG.log
timeout 900 python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err
echo "bench exit $?"; cut -c1-300 $O/bench_${TAG}_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err
timeout 600 python bench.py --config infer > $O/bench_${TAG}_infer.json 2> $O/bench_${TAG}_infer.err
echo "infer exit $?"; cut -c1-200 $O/bench_${TAG}_infer.json
