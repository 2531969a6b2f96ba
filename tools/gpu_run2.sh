set -x
cd $GRAFT_REPO_ROOT
for m in 1 0 2; do
  SPEAR_BATCH_GIANT_MULTI=$m timeout 600 python bench.py --steps 5 --warmup 3 --no-token --no-tuned --no-cpu-baseline > gpurun_out/r2_b2_multi$m.json 2> gpurun_out/r2_b2_multi$m.err
done
timeout 600 python tools/profile_step.py --steps 3 --classes > gpurun_out/r2_classes_v1.log 2>&1 &&
S=$(python tools/profile_step.py --count-only) &&
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 200 --csv --log-file gpurun_out/r2_v1_launches.csv python tools/profile_step.py > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_intt_modup_fwd_a|k_ntt_b_ks_all|k_pmac_tma|k_ks_baby_fused' -s 4 -c 4 -o gpurun_out/r2_v1_full python tools/profile_step.py > gpurun_out/r2_ncu_b.log 2>&1
ls -la gpurun_out | tail -8
