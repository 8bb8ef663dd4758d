#!/bin/bash
# Last profiling pass of round 2 (PK_DESC2 with compacted descriptors and converted column words; run under gpurun on one B200): launch list of the bench command, then one
# `ncu --set full` capture of the Gram tile launch for the tiers whose kernel changed.  Every ncu run follows a plain run of
# the same command line.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --e2e-reps 1"
$B > $O/r02e_launchlist_plain.json 2> $O/r02e_launchlist_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02e_launches_bench.csv $B > $O/r02e_launchlist_ncu.log 2>&1
echo "launch list rc=$?"
scripts/profile_one.sh r02e_500tips 192 500 64 0:0:0
scripts/profile_one.sh r02e_100tips 1024 100 64 0:0:0
scripts/profile_one.sh r02e_500tips_fp32 192 500 32 0:0:0
