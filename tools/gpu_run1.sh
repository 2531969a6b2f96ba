set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1
nproc
SPEAR_FUSED_MODUP=0 timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py -q -m gpu > gpurun_out/r2_t1_fullsize_unfused.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1_fullsize_unfused.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2_t1_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "rc=$?" >> gpurun_out/r2_b1.err
tail -3 gpurun_out/r2_t1_fullsize_unfused.log gpurun_out/r2_t1_all.log gpurun_out/r2_b1.err
