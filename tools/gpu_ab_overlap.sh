# A/B: the diagonal MAC of row chunk i next to the baby-step key stream of chunk i+1 (SPEAR_PIPE_ROWS), with the baby kernel's
# residency capped (SPEAR_BABY_SMEM_PAD) so that MAC CTAs fit beside it; reference split and the hoisting-aware split
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
for w in 1 8; do
  for cfg in "0 0" "3 0" "3 43000" "3 81000" "9 43000" "9 81000" "0 43000"; do
    set -- $cfg
    echo "== weight $w pipe $1 pad $2" >> $O/r2_ab_overlap.log
    SPEAR_PIPE_ROWS=$1 SPEAR_BABY_SMEM_PAD=$2 timeout 200 python tools/profile_step.py --steps 3 --classes --weight $w 2>&1 | grep -E "per mat-vec|pmac|ks_baby|max_abs" >> $O/r2_ab_overlap.log
  done
done
cat $O/r2_ab_overlap.log
