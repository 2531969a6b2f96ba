set -x
cd $GRAFT_REPO_ROOT
timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1_v6.json 2> gpurun_out/r2_c5_n1_v6.err
timeout 900 python tools/fully_enc_bench.py --phases > gpurun_out/r2_c5_n1_v6_phases.json 2> gpurun_out/r2_c5_n1_v6_phases.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tuned --no-token > gpurun_out/r2_b12.json 2> gpurun_out/r2_b12.err; tail -3 gpurun_out/r2_b12.err
