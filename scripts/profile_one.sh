#!/bin/bash
# One `ncu --set full --import-source on` capture of the Gram tile launch of a probe (after a plain run of the same command);
# exports the raw and source pages as CSV into gpurun_out/.  usage: scripts/profile_one.sh <name> <probe args...>
set -u
O=gpurun_out
name=$1; shift
python scripts/probe.py "$@" > $O/probe_$name.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pk_dp_kernel -s 1 -c 1 -f -o $O/prof_$name python scripts/probe.py "$@" > $O/ncu_$name.log 2>&1
echo "capture $name rc=$?"; tail -1 $O/probe_$name.log
ncu -i $O/prof_$name.ncu-rep --page raw --csv > $O/prof_$name.raw.csv 2>/dev/null
ncu -i $O/prof_$name.ncu-rep --page source --csv > $O/prof_$name.source.csv 2>/dev/null
rm -f $O/prof_$name.ncu-rep
