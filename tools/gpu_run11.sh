set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/client_bench.py > gpurun_out/r2_client_legs.json 2> gpurun_out/r2_client_legs.err
SPEAR_FUSED_CLIENT=0 timeout 300 python tools/client_bench.py > gpurun_out/r2_client_legs_staged.json 2> gpurun_out/r2_client_legs_staged.err
timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1_v5.json 2> gpurun_out/r2_c5_n1_v5.err
cat gpurun_out/r2_client_legs.json gpurun_out/r2_client_legs_staged.json
