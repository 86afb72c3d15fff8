#!/bin/bash
# Multi-GPU evidence in one gpurun call: peer-memory probe + fused all-reduce/Adam correctness and timing, then the bench
# at N ranks (fused peer all-reduce and the NCCL variant).   usage: bash profiles/gpu_dp_session.sh <tag> <N>
TAG=${1:-r1_dp}
N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for be in symm; do
  timeout 300 $TR --master-port 29511 tests/dp_peer_check.py --backend $be > $O/dp_check_${TAG}_$be.json 2> $O/dp_check_${TAG}_$be.err
  echo "dp_check $be exit $?"; tail -1 $O/dp_check_${TAG}_$be.json | cut -c1-400; tail -3 $O/dp_check_${TAG}_$be.err | cut -c1-300
done
timeout 600 $TR --master-port 29512 bench.py --gpus $N --no-cpu-baseline --no-ref-cuda > $O/bench_${TAG}_n$N.json 2> $O/bench_${TAG}_n$N.err
echo "bench N=$N exit $?"; cut -c1-300 $O/bench_${TAG}_n$N.json; tail -3 $O/bench_${TAG}_n$N.err | cut -c1-300
timeout 600 $TR --master-port 29513 bench.py --gpus $N --nccl --no-cpu-baseline --no-ref-cuda > $O/bench_${TAG}_n${N}_nccl.json 2> $O/bench_${TAG}_n${N}_nccl.err
echo "bench nccl N=$N exit $?"; cut -c1-300 $O/bench_${TAG}_n${N}_nccl.json
timeout 300 python bench.py --no-cpu-baseline --no-ref-cuda > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err
echo "bench N=1 exit $?"; cut -c1-300 $O/bench_${TAG}_n1.json
