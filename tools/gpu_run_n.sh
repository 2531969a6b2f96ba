# usage: gpurun --gpus N -- 'bash tools/gpu_run_n.sh N [extra bench flags]'
set -x
cd $GRAFT_REPO_ROOT
N=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2_bench_n$N.err
tail -n 5 gpurun_out/r2_bench_n$N.err
