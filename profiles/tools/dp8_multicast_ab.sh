O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tests/dp_peer_check.py > $O/dp_check_r1_dp8c_mc.json 2> $O/dp_check_r1_dp8c_mc.err; echo "mc exit $?"; tail -1 $O/dp_check_r1_dp8c_mc.json | cut -c1-420
NGP_DP_MULTICAST=0 timeout 200 $TR --master-port 29512 tests/dp_peer_check.py > $O/dp_check_r1_dp8c_p2p.json 2> $O/dp_check_r1_dp8c_p2p.err; echo "p2p exit $?"; tail -1 $O/dp_check_r1_dp8c_p2p.json | cut -c1-420
NGP_DP_MULTICAST=0 timeout 300 $TR --master-port 29513 bench.py --gpus 8 --no-cpu-baseline --no-ref-cuda > $O/bench_r1_dp8c_n8_p2p.json 2> $O/bench_r1_dp8c_n8_p2p.err; echo "bench p2p exit $?"; cut -c1-240 $O/bench_r1_dp8c_n8_p2p.json
timeout 300 $TR --master-port 29514 bench.py --gpus 8 --no-cpu-baseline --no-ref-cuda > $O/bench_r1_dp8c_n8_mc.json 2> $O/bench_r1_dp8c_n8_mc.err; echo "bench mc exit $?"; cut -c1-240 $O/bench_r1_dp8c_n8_mc.json
