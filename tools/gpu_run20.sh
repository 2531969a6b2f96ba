set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_ntt_b_ks_all|k_sum_groups|k_moddown' -s 2 -c 4 -o $O/r2_s3_ks python tools/profile_step.py > $O/r2_ncu20.log 2>&1
tail -n 3 $O/r2_ncu20.log
