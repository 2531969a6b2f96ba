# Final single-GPU evidence pass of round 2 (run under gpurun; outputs in gpurun_out/, copied to profiles/ afterwards)
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python __graft_entry__.py smoke > $O/r2_final_smoke.log 2>&1; echo "rc=$?" >> $O/r2_final_smoke.log
timeout 1800 python -m pytest tests -q -m gpu > $O/r2_final_gputest.log 2>&1; echo "rc=$?" >> $O/r2_final_gputest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err; echo "rc=$?" >> $O/r2_final_bench_n1.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2_final_bench_reference_n1.json 2> $O/r2_final_bench_reference_n1.err
timeout 900 python bench.py --config c2 --steps 20 --warmup 5 --no-token > $O/r2_final_bench_c2.json 2> $O/r2_final_bench_c2.err
timeout 600 python tools/fully_enc_bench.py > $O/r2_final_c5_n1.json 2> $O/r2_final_c5_n1.err
timeout 300 python tools/client_bench.py > $O/r2_final_client_legs.json 2> $O/r2_final_client_legs.err
timeout 300 python tools/profile_step.py --steps 3 --classes > $O/r2_final_classes.log 2>&1
# ncu: launch list of one un-overlapped mat-vec and of bench.py's own timed steps; full-set capture of the four big kernels
S=$(python tools/profile_step.py --count-only) &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 200 --csv --log-file $O/r2_final_matvec_launches.csv python tools/profile_step.py > $O/r2_final_ncu_a.log 2>&1
FLAGS="--steps 2 --warmup 3 --no-cpu-baseline --no-tuned --no-token"
S2=$(python bench.py $FLAGS --count-only) &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $S2 -c 260 --csv --log-file $O/r2_final_bench_launches.csv python bench.py $FLAGS > $O/r2_final_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_intt_modup_fwd_a|k_ntt_b_ks_all|k_pmac_tma|k_ks_baby_fused' -s 5 -c 5 -o $O/r2_final_full python tools/profile_step.py > $O/r2_final_ncu_c.log 2>&1
tail -n 3 $O/r2_final_smoke.log $O/r2_final_gputest.log $O/r2_final_bench_n1.err
