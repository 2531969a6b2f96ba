set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sharded_latency.py > gpurun_out/r2_lat_n2.json 2> gpurun_out/r2_lat_n2.err
SPEAR_FUSED_MODUP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_latency.py > gpurun_out/r2_lat_n2_unfused.json 2> gpurun_out/r2_lat_n2_unfused.err
SPEAR_PEER=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 --no-token --no-cpu-baseline > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
cat gpurun_out/r2_lat_n2.json gpurun_out/r2_lat_n2_unfused.json
