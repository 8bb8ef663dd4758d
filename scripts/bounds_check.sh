#!/bin/bash
# compute-sanitizer is not available on the GPU pool: this builds the library with -DPK_DEBUG_BOUNDS (every C-cell index
# checked against the slab / shared-memory array, violations counted, the call fails) and runs the mixed small workload
# and the GPU parity tests against that build.  Run under gpurun.
set -euo pipefail
cd "$(dirname "$0")/.."
PK_DEFS="-DPK_DEBUG_BOUNDS" PK_OUT="$PWD/kamphir_b200/libphylok_dbg.so" bash kamphir_b200/csrc/build.sh
# every step under its own timeout: a deadlock found this way must not eat the GPU box's time limit
PK_LIB="$PWD/kamphir_b200/libphylok_dbg.so" timeout 120 python scripts/sanitize_case.py
PK_LIB="$PWD/kamphir_b200/libphylok_dbg.so" timeout 300 python -m pytest tests -m gpu -x -q
PK_LIB="$PWD/kamphir_b200/libphylok_dbg.so" timeout 120 python scripts/fuzz_parity.py 60
