set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bsgs_paths.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2_t10.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t10.log
for f in 1 0; do
  SPEAR_FUSED_ENCODE=$f timeout 300 python tools/encode_bench.py > gpurun_out/r2_encode_c5_fused$f.json 2> gpurun_out/r2_encode_c5_fused$f.err
  SPEAR_FUSED_ENCODE=$f timeout 300 python tools/encode_bench.py --N 32768 --L0 24 > gpurun_out/r2_encode_c3_fused$f.json 2> gpurun_out/r2_encode_c3_fused$f.err
done
timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1_v4.json 2> gpurun_out/r2_c5_n1_v4.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err
SPEAR_FUSED_CLIENT=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tuned > gpurun_out/r2_b10_staged_client.json 2> gpurun_out/r2_b10_staged_client.err
tail -n 3 gpurun_out/r2_t10.log; cat gpurun_out/r2_encode_*.json
