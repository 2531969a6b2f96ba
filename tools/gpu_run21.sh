set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bsgs_paths.py -q -m gpu -k "two_phase" -x > $O/r2_t21.log 2>&1; echo "rc=$?" >> $O/r2_t21.log
timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py -q -m gpu -k "c3" -x >> $O/r2_t21.log 2>&1; echo "rc=$?" >> $O/r2_t21.log
tail -n 12 $O/r2_t21.log
