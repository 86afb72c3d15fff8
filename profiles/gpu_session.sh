#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, smoke, the bench lines, the ncu launch list and one
# `ncu --set full` capture of the hot kernels.  Everything lands in gpurun_out/ (scratch); profiles/summarize.py
# turns the ncu outputs into the summaries committed under profiles/.
#   usage: bash profiles/gpu_session.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-r1_v7}
DO_TESTS=${2:-tests}
DO_NCU=${3:-ncu}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/gpu_$TAG.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv >> $O/gpu_$TAG.txt 2>&1
if [ "$DO_TESTS" = tests ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
  tail -5 $O/pytest_gpu_$TAG.log
  timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1
  echo "smoke exit $?" >> $O/smoke_$TAG.log
  tail -2 $O/smoke_$TAG.log
fi
timeout 900 python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err
echo "bench exit $?"
cut -c1-600 $O/bench_${TAG}_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err
timeout 600 python bench.py --no-cpu-baseline --no-ref-cuda --kernel-table $O/ktable_$TAG.txt > $O/bench_${TAG}_kt.json 2> $O/bench_${TAG}_kt.err
for v in 1 2 4; do
  timeout 600 python bench.py --views $v --no-cpu-baseline --no-ref-cuda --steps 96 --warmup 16 > $O/bench_${TAG}_views$v.json 2> $O/bench_${TAG}_views$v.err
  cut -c1-300 $O/bench_${TAG}_views$v.json
done
if [ "$DO_NCU" = ncu ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --profile-steps 1 > $O/ncu_launches_$TAG.log 2>&1
  echo "ncu launches exit $?"
  timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'field_forward|field_backward|encode_backward_warpagg|scatter|march_slab|march_count|march_write|march_compact|composite_train|adam_step' \
    --launch-skip 60 -c 14 -f -o $O/prof_$TAG \
    python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --profile-steps 1 > $O/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?"
  ls -la $O/prof_$TAG.ncu-rep
fi
