set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bsgs_paths.py -q -m gpu -k "two_phase or giant_sharding" -x > $O/r2_t14.log 2>&1; echo "rc=$?" >> $O/r2_t14.log
timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py -q -m gpu -k "c3" -x >> $O/r2_t14.log 2>&1; echo "rc=$?" >> $O/r2_t14.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-token --no-tuned --no-cpu-baseline > $O/r2_b14.json 2> $O/r2_b14.err; echo "rc=$?" >> $O/r2_b14.err
tail -n 30 $O/r2_t14.log
