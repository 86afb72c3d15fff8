#!/bin/bash
# The short version of gpu_session.sh (about 3 GPU-minutes): parity tests, smoke, the two bench arms, the ncu launch list
# of the graphed step and one `ncu --set full` capture of every hot kernel of one step (single ray chain, so each launch
# sees the whole 3.4 M-sample step).      usage: bash profiles/gpu_session_slim.sh <tag>
TAG=${1:-r1_v8}
O=gpurun_out
mkdir -p $O
( time timeout 120 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 90 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $O/smoke_$TAG.log; tail -2 $O/smoke_$TAG.log
timeout 200 python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err
echo "bench exit $?"; cut -c1-330 $O/bench_${TAG}_default.json; tail -2 $O/bench_${TAG}_default.err | cut -c1-300
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --profile-steps 1 > $O/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
timeout 150 ncu --set full --clock-control none --import-source on \
  -k regex:'field_forward|field_backward|encode_backward_warpagg|march_slab|march_scan|march_compact|train_ray_loss|adam_step_fused' \
  --launch-skip 42 -c 9 -f -o $O/prof_$TAG \
  python bench.py --no-graph --chunks 1 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --profile-steps 1 > $O/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"; ls -la $O/prof_$TAG.ncu-rep
