#!/bin/bash
# Evidence pass on one B200 (run under gpurun): GPU test suite, bench line, ncu launch lists of bench.py's timed
# steps and of one mat-vec, one --set full capture of the main kernels.  Everything lands in gpurun_out/$TAG_*.
TAG=${1:-r1s2}
O=gpurun_out
mkdir -p $O
[ -z "$SKIP_PYTEST" ] && python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/${TAG}_pytest_gpu.log
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err || exit 1
[ -z "$SKIP_REFERENCE" ] && python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
# launch list of bench.py's own timed steps (2 steps x 3 projections), skipping set-up and warm-up launches
# (SKIP_BENCH_NCU=1 leaves this pass out: ncu walks the ~4000 set-up launches of bench.py one by one, ~7 minutes)
if [ -z "$SKIP_BENCH_NCU" ]; then
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-tuned"
S=$(python bench.py $BA --count-only | tail -1)
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 1400 --csv --log-file $O/${TAG}_bench_launches.csv \
    python bench.py $BA > $O/${TAG}_bench_under_ncu.log 2>&1
fi
# one un-overlapped mat-vec
S1=$(python tools/profile_step.py --count-only | tail -1)
ncu --metrics gpu__time_duration.sum --clock-control none -s $S1 -c 400 --csv --log-file $O/${TAG}_matvec_launches.csv \
    python tools/profile_step.py > $O/${TAG}_matvec_under_ncu.log 2>&1
python tools/profile_step.py --classes --steps 5 --warmup 2 > $O/${TAG}_classes.log 2>&1
# full-set capture, one launch of each main kernel of the profiled mat-vec
ncu --set full --clock-control none --import-source on \
    -k regex:'k_ks_baby_fused|k_ntt_b_ks|k_pmac_tma|ntt_fwd_a2|k_modup|k_sum_groups' --launch-skip-before-match $S1 -c 9 \
    -o $O/${TAG}_full python tools/profile_step.py > $O/${TAG}_full_under_ncu.log 2>&1
ls -la $O | grep $TAG
