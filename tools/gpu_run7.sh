set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_reference_scripts.py -q -m gpu > gpurun_out/r2_t7_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t7_ref.log
timeout 900 python tools/fully_enc_bench.py --phases > gpurun_out/r2_c5_n1_phases.json 2> gpurun_out/r2_c5_n1_phases.err
timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1.json 2> gpurun_out/r2_c5_n1.err
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-token > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
tail -n 15 gpurun_out/r2_t7_ref.log
