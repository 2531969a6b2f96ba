set -x
cd $GRAFT_REPO_ROOT
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_bsgs_paths.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r2_t13.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t13.log
for f in 1 0; do
  SPEAR_ENCODE_TILED=$f timeout 300 python tools/encode_bench.py > gpurun_out/r2_encode_c5_tiled$f.json 2> gpurun_out/r2_encode_c5_tiled$f.err
  SPEAR_ENCODE_TILED=$f timeout 300 python tools/encode_bench.py --N 32768 --L0 24 > gpurun_out/r2_encode_c3_tiled$f.json 2> gpurun_out/r2_encode_c3_tiled$f.err
done
timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1_v7.json 2> gpurun_out/r2_c5_n1_v7.err
./tools/ubench/exchange > gpurun_out/r2_ubench_exchange.log 2>&1
tail -n 3 gpurun_out/r2_smoke.log gpurun_out/r2_t13.log; cat gpurun_out/r2_encode_*tiled*.json gpurun_out/r2_ubench_exchange.log
