#!/bin/bash
# Developer helper: A/B of kernel libraries over the CTA tiers (fp64 Gram probes).  usage: scripts/ab_tiers.sh <out file> lib1 lib2 ...
out=$1; shift
: > $out
for spec in "512 500" "1024 100" "512 200" "256 800" "192 1000" "128 1300" "128 1500" "96 2000"; do
  for v in "$@"; do
    lib=kamphir_b200/var/lib_$v.so; [ "$v" = main ] && lib=kamphir_b200/libphylok.so
    echo "== $v: probe $spec" >> $out
    PK_LIB=$lib timeout 300 python scripts/probe.py $spec 64 0:0:0 2>&1 | grep -v "^gen" >> $out
  done
done
