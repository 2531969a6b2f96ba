# Multi-GPU evidence of round 2 (run under gpurun --gpus N; outputs in gpurun_out/): strong-scaling bench with the measured
# token loop, config 5 (fully encrypted FFN blocks) sharded over N GPUs, and at N = 2 the real-window GPU tests
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
N=${1:-2}
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_peer_exchange.py -q -m gpu > $O/r2_final_peer_tests_n2.log 2>&1; echo "rc=$?" >> $O/r2_final_peer_tests_n2.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_final_bench_n$N.json 2> $O/r2_final_bench_n$N.err; echo "rc=$?" >> $O/r2_final_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/fully_enc_bench.py > $O/r2_final_c5_n$N.json 2> $O/r2_final_c5_n$N.err; echo "rc=$?" >> $O/r2_final_c5_n$N.err
tail -n 3 $O/r2_final_bench_n$N.err $O/r2_final_c5_n$N.err
