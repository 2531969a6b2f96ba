set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_b16_n$N.json 2> $O/r2_b16_n$N.err; echo "rc=$?" >> $O/r2_b16_n$N.err
tail -n 5 $O/r2_b16_n$N.err
