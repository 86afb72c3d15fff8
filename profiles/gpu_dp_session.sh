#!/bin/bash
# Multi-GPU evidence in one gpurun call: fused peer all-reduce/Adam + sharded occupancy refresh correctness and timing
# (tests/dp_peer_check.py), then the bench at the listed rank counts (fused peer all-reduce; NCCL variant at the largest).
#   usage: bash profiles/gpu_dp_session.sh <tag> "<N list, e.g. 8 4 2 1>"
TAG=${1:-r1_dp}
NS=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
FIRST=1
for N in $NS; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  if [ "$N" = 1 ]; then TR="python"; fi
  if [ "$FIRST" = 1 ] && [ "$N" != 1 ]; then
    timeout 300 $TR --master-port 29511 tests/dp_peer_check.py > $O/dp_check_${TAG}_n$N.json 2> $O/dp_check_${TAG}_n$N.err
    echo "dp_check N=$N exit $?"; tail -1 $O/dp_check_${TAG}_n$N.json | cut -c1-500; grep -v "^\*\*\*\|OMP_NUM\|^$" $O/dp_check_${TAG}_n$N.err | tail -3 | cut -c1-300
    NGP_DP_MULTICAST=0 timeout 300 $TR --master-port 29514 tests/dp_peer_check.py > $O/dp_check_${TAG}_n${N}_p2p.json 2> $O/dp_check_${TAG}_n${N}_p2p.err
    echo "dp_check (P2P only) N=$N exit $?"; tail -1 $O/dp_check_${TAG}_n${N}_p2p.json | cut -c1-500
  fi
  PORT=$((29520 + N))
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --no-cpu-baseline --no-ref-cuda > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
  else
    timeout 600 $TR --master-port $PORT bench.py --gpus $N --no-cpu-baseline --no-ref-cuda > $O/bench_${TAG}_n$N.json 2> $O/bench_${TAG}_n$N.err
  fi
  echo "bench N=$N exit $?"; cut -c1-260 $O/bench_${TAG}_n$N.json; grep -v "^\*\*\*\|OMP_NUM\|^$" $O/bench_${TAG}_n$N.err | tail -2 | cut -c1-300
  if [ "$FIRST" = 1 ] && [ "$N" != 1 ]; then
    timeout 600 $TR --master-port $((PORT + 20)) bench.py --gpus $N --nccl --no-cpu-baseline --no-ref-cuda > $O/bench_${TAG}_n${N}_nccl.json 2> $O/bench_${TAG}_n${N}_nccl.err
    echo "bench nccl N=$N exit $?"; cut -c1-260 $O/bench_${TAG}_n${N}_nccl.json
  fi
  FIRST=0
done
