#!/bin/bash
# Developer helper: time kernel variants.  usage: scripts/variants.sh "<probe args>" name[:ENV=VAL,...] ...
# name = a library kamphir_b200/var/lib_<name>.so (or "main" for kamphir_b200/libphylok.so); optional environment settings
# after a colon (e.g. main:PK_STRIP=2).
args=$1; shift
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  lib=kamphir_b200/var/lib_$v.so; [ "$v" = main ] && lib=kamphir_b200/libphylok.so
  echo "== $spec: probe $args"
  env $envs PK_LIB=$lib python scripts/probe.py $args 2>&1 | grep -v "^gen"
done
