set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python __graft_entry__.py smoke > $O/r2_t17.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu -x >> $O/r2_t17.log 2>&1; echo "rc=$?" >> $O/r2_t17.log
F="--steps 10 --warmup 3 --no-token --no-tuned --no-cpu-baseline"
timeout 300 python bench.py $F > $O/r2_b17.json 2> $O/r2_b17.err
SPEAR_FUSED_FINISH=0 timeout 300 python bench.py $F > $O/r2_b17_nofinish.json 2>> $O/r2_b17.err
for k in 2 3 4 9; do SPEAR_PIPE_ROWS=$k timeout 300 python bench.py $F > $O/r2_b17_pipe$k.json 2>> $O/r2_b17.err; done
tail -n 8 $O/r2_t17.log
