#!/bin/bash
# Round-2 final evidence session (1 GPU): GPU tests, smoke, the bench lines (default, reference arm, infer, encoder, the
# pipelined-optimizer A/B), the ncu launch list of the default bench command and one `ncu --set full` capture of the hot
# kernels (eager steps, single chain).  Everything lands in gpurun_out/.
#   usage: bash profiles/evidence_r2_final.sh <tag> [tests|notests] [full|nofull]
TAG=${1:-r2_final}
DO_TESTS=${2:-tests}
DO_FULL=${3:-full}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/gpu_$TAG.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv >> $O/gpu_$TAG.txt 2>&1
if [ "$DO_TESTS" = tests ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
  tail -4 $O/pytest_gpu_$TAG.log
  timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1
  echo "smoke exit $?" >> $O/smoke_$TAG.log
  tail -2 $O/smoke_$TAG.log
fi
timeout 900 python bench.py > $O/bench_${TAG}_default.json 2> $O/bench_${TAG}_default.err
echo "bench exit $?"; cut -c1-300 $O/bench_${TAG}_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err
timeout 600 python bench.py --config infer > $O/bench_${TAG}_infer.json 2> $O/bench_${TAG}_infer.err
echo "infer exit $?"; cut -c1-200 $O/bench_${TAG}_infer.json
timeout 600 python bench.py --config encoder > $O/bench_${TAG}_encoder.json 2> $O/bench_${TAG}_encoder.err
echo "encoder exit $?"; cut -c1-200 $O/bench_${TAG}_encoder.json
LITE="--no-cpu-baseline --no-ref-cuda --no-shading --profile-steps 2"
timeout 300 python bench.py --pipeline $LITE > $O/bench_${TAG}_n1_pipelined.json 2> $O/bench_${TAG}_n1_pipelined.err
echo "pipelined exit $?"; cut -c1-200 $O/bench_${TAG}_n1_pipelined.json
timeout 300 python bench.py --views 1 --steps 200 --warmup 20 $LITE --timeline $O/timeline_${TAG}_v1.json > $O/bench_${TAG}_views1.json 2> $O/bench_${TAG}_views1.err
timeout 300 python bench.py --steps 20 --warmup 5 $LITE --timeline $O/timeline_${TAG}_v8.json > /dev/null 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-shading --profile-steps 1"
timeout 600 $CMD > $O/ncu_plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
if [ "$DO_FULL" = full ]; then
  CMD2="python bench.py --no-graph --chunks 1 --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-shading --profile-steps 1"
  timeout 600 $CMD2 > $O/ncu_plain2_$TAG.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'field_forward|field_backward|encode_backward_warpagg|march_packed|train_ray_loss|adam_step|quad_table|check_finite' \
    --launch-skip 40 -c 14 -f -o $O/prof_$TAG $CMD2 > $O/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?"
  ls -la $O/prof_$TAG.ncu-rep
fi
