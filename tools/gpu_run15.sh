set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -q -m gpu -x > $O/r2_t15.log 2>&1; echo "rc=$?" >> $O/r2_t15.log
timeout 600 python -m pytest tests/test_gpu_bsgs_paths.py -q -m gpu -k "two_phase" -x >> $O/r2_t15.log 2>&1; echo "rc=$?" >> $O/r2_t15.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_b15_n$N.json 2> $O/r2_b15_n$N.err; echo "rc=$?" >> $O/r2_b15_n$N.err
SPEAR_TWO_PHASE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_b15_n${N}_old.json 2> $O/r2_b15_n${N}_old.err; echo "rc=$?" >> $O/r2_b15_n${N}_old.err
tail -n 30 $O/r2_t15.log; tail -n 5 $O/r2_b15_n$N.err
