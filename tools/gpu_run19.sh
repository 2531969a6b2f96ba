set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python __graft_entry__.py smoke > $O/r2_t19.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_bsgs_paths.py tests/test_gpu_fullsize_parity.py tests/test_gpu_parity.py -q -m gpu -x >> $O/r2_t19.log 2>&1; echo "rc=$?" >> $O/r2_t19.log
F="--steps 10 --warmup 3 --no-token --no-cpu-baseline"
timeout 300 python bench.py $F > $O/r2_b19.json 2> $O/r2_b19.err
timeout 300 python tools/profile_step.py --steps 3 --classes > $O/r2_classes19.log 2>&1
tail -n 6 $O/r2_t19.log
