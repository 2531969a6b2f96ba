set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_reference_scripts.py -q -m gpu -x > gpurun_out/r2_t6_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t6_ref.log
tail -n 40 gpurun_out/r2_t6_ref.log
