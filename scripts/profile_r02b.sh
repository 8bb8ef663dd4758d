#!/bin/bash
# Second profiling pass of round 2 (pk_rect2 kernels; run under gpurun on one B200): launch list of the bench command, then one
# `ncu --set full` capture of the Gram tile launch per kernel tier.  Every ncu run follows a plain run of the same command line.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --e2e-reps 1"
$B > $O/r02b_launchlist_plain.json 2> $O/r02b_launchlist_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_bench.csv $B > $O/r02b_launchlist_ncu.log 2>&1
echo "launch list rc=$?"
scripts/profile_one.sh r02b_500tips 192 500 64 0:0:0
scripts/profile_one.sh r02b_100tips 1024 100 64 0:0:0
scripts/profile_one.sh r02b_2000tips 96 2000 64 0:0:0
scripts/profile_one.sh r02b_deal_s3 192 500 64 0:0:0 3
scripts/profile_one.sh r02b_500tips_fp32 192 500 32 0:0:0
