set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -q -m gpu > gpurun_out/r2_t5_peer.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t5_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sharded_latency.py > gpurun_out/r2_lat_n2.json 2> gpurun_out/r2_lat_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tools/sharded_latency.py --weight 1 > gpurun_out/r2_lat_n2_w1.json 2> gpurun_out/r2_lat_n2_w1.err
bash tools/gpu_run_n.sh 2
tail -n 2 gpurun_out/r2_t5_peer.log; cat gpurun_out/r2_lat_n2.json gpurun_out/r2_lat_n2_w1.json
