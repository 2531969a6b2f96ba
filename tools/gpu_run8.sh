set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_reference_scripts.py > gpurun_out/r2_t8_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t8_all.log
for f in 1 0; do
  SPEAR_FUSED_ENCODE=$f timeout 300 python tools/encode_bench.py > gpurun_out/r2_encode_c5_fused$f.json 2> gpurun_out/r2_encode_c5_fused$f.err
  SPEAR_FUSED_ENCODE=$f timeout 300 python tools/encode_bench.py --N 32768 --L0 24 > gpurun_out/r2_encode_c3_fused$f.json 2> gpurun_out/r2_encode_c3_fused$f.err
done
timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1_v2.json 2> gpurun_out/r2_c5_n1_v2.err
tail -n 4 gpurun_out/r2_t8_all.log; cat gpurun_out/r2_encode_*.json
