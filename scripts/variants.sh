#!/bin/bash
# Developer helper: time the kernel variants built into kamphir_b200/var/lib_<name>.so (PK_LIB selects the library).
# usage: scripts/variants.sh "<probe args>" name...
args=$1; shift
for v in "$@"; do
  echo "== $v: probe $args"
  PK_LIB=kamphir_b200/var/lib_$v.so python scripts/probe.py $args 2>&1 | grep -v "^gen"
done
