#!/bin/bash
# Round-2 profiling pass (run under gpurun on one B200): launch list of the bench command, then one `ncu --set full` capture
# of the Gram tile launch per kernel tier the judge asked for.  Every ncu run follows a plain run of the same command line.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --e2e-reps 1"
$B > $O/r02_launchlist_plain.json 2> $O/r02_launchlist_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv $B > $O/r02_launchlist_ncu.log 2>&1
echo "launch list rc=$?"
cap() {   # name, probe args
  local name=$1; shift
  python scripts/probe.py "$@" > $O/r02_probe_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pk_dp_kernel -s 1 -c 1 -f -o $O/prof_r02_$name python scripts/probe.py "$@" > $O/r02_ncu_$name.log 2>&1
  echo "capture $name rc=$?"; tail -1 $O/r02_probe_$name.log
  # the report is ~23 MB with sources imported and gpurun brings back 64 MiB at most: export the two pages, drop the report
  ncu -i $O/prof_r02_$name.ncu-rep --page raw --csv > $O/prof_r02_$name.raw.csv 2>/dev/null
  ncu -i $O/prof_r02_$name.ncu-rep --page source --csv > $O/prof_r02_$name.source.csv 2>/dev/null
  rm -f $O/prof_r02_$name.ncu-rep
}
cap 500tips 192 500 64 0:0:0
cap 100tips 1024 100 64 0:0:0
cap 2000tips 96 2000 64 0:0:0
cap deal_s3 192 500 64 0:0:0 3
cap 500tips_fp32 192 500 32 0:0:0
