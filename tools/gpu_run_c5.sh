# usage: gpurun --gpus N -- 'bash tools/gpu_run_c5.sh N'
set -x
cd $GRAFT_REPO_ROOT
N=$1
if [ "$N" = "1" ]; then
  timeout 900 python tools/fully_enc_bench.py --phases > gpurun_out/r2_c5_n1_phases.json 2> gpurun_out/r2_c5_n1_phases.err
  timeout 900 python tools/fully_enc_bench.py > gpurun_out/r2_c5_n1.json 2> gpurun_out/r2_c5_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/fully_enc_bench.py > gpurun_out/r2_c5_n$N.json 2> gpurun_out/r2_c5_n$N.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/fully_enc_bench.py --phases > gpurun_out/r2_c5_n${N}_phases.json 2> gpurun_out/r2_c5_n${N}_phases.err
fi
tail -n 3 gpurun_out/r2_c5_n$N.err
