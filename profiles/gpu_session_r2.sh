#!/bin/bash
# Round-2 GPU session: GPU parity tests (all, no -x), smoke, the default bench line, the reference arm, and - optionally -
# the ncu launch list of the same bench command.  Everything lands in gpurun_out/ (scratch).
#   usage: bash profiles/gpu_session_r2.sh <tag> [tests|notests] [ncu|noncu] [pytest -k expression]
TAG=${1:-r2_v1}
DO_TESTS=${2:-tests}
DO_NCU=${3:-noncu}
KEXPR=${4:-}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/gpu_$TAG.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv >> $O/gpu_$TAG.txt 2>&1
if [ "$DO_TESTS" = tests ]; then
  if [ -n "$KEXPR" ]; then
    ( time timeout 1500 python -m pytest tests -m gpu -q -k "$KEXPR" ) > $O/pytest_gpu_$TAG.log 2>&1
  else
    ( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu_$TAG.log 2>&1
  fi
  echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
  tail -15 $O/pytest_gpu_$TAG.log
  timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1
  echo "smoke exit $?" >> $O/smoke_$TAG.log
  tail -2 $O/smoke_$TAG.log
fi
timeout 900 python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err
echo "bench exit $?"
cut -c1-400 $O/bench_${TAG}_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err
if [ "$DO_NCU" = ncu ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --profile-steps 1 > $O/ncu_plain_$TAG.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --profile-steps 1 > $O/ncu_launches_$TAG.log 2>&1
  echo "ncu launches exit $?"
fi
