set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python __graft_entry__.py smoke > $O/r2_final_smoke.log 2>&1; echo "rc=$?" >> $O/r2_final_smoke.log
timeout 1800 python -m pytest tests -q -m gpu > $O/r2_final_gputest.log 2>&1; echo "rc=$?" >> $O/r2_final_gputest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err; echo "rc=$?" >> $O/r2_final_bench_n1.err
tail -n 3 $O/r2_final_smoke.log $O/r2_final_gputest.log $O/r2_final_bench_n1.err
