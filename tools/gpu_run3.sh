set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -q -m gpu > gpurun_out/r2_t3_peer.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t3_peer.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_b3_n2.json 2> gpurun_out/r2_b3_n2.err; echo "rc=$?" >> gpurun_out/r2_b3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2_b3_ref_n2.json 2> gpurun_out/r2_b3_ref_n2.err
tail -n 3 gpurun_out/r2_t3_peer.log gpurun_out/r2_b3_n2.err
