#!/bin/bash
out=$1; shift
: > $out
for spec in "512 500" "1024 100" "512 200" "256 800" "192 1000" "128 1300" "96 2000"; do
  for v in "$@"; do
    lib=kamphir_b200/var/lib_$v.so
    echo "== $v: probe $spec" >> $out
    PK_LIB=$lib timeout 300 python scripts/probe.py $spec 32 0:0:0 2>&1 | grep -v "^gen" >> $out
  done
done
