set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_bsgs_paths.py -q -m gpu -k "host_buffer or two_phase" -x > $O/r2_t18.log 2>&1; echo "rc=$?" >> $O/r2_t18.log
F="--steps 10 --warmup 3 --no-token --no-tuned --no-cpu-baseline"
timeout 300 python bench.py $F > $O/r2_b18.json 2> $O/r2_b18.err
tail -n 8 $O/r2_t18.log
