# quick 2-GPU check: smoke, two-phase bench, and the fallback plan when peer windows are switched off
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python __graft_entry__.py smoke > $O/r2_check_smoke.log 2>&1; echo "rc=$?" >> $O/r2_check_smoke.log
F="--gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --tokens 1"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py $F > $O/r2_check_n2.json 2> $O/r2_check_n2.err; echo "rc=$?" >> $O/r2_check_n2.err
SPEAR_PEER=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py $F > $O/r2_check_n2_nopeer.json 2> $O/r2_check_n2_nopeer.err; echo "rc=$?" >> $O/r2_check_n2_nopeer.err
tail -n 2 $O/r2_check_smoke.log $O/r2_check_n2.err $O/r2_check_n2_nopeer.err
