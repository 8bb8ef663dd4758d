#!/bin/bash
# Developer helper: builds one kernel tier only (-DPK_DEV_TIER=<threads>, unlabelled fp64) into /tmp/lib_dev_<tag>.so, dumps its
# SASS and prints registers / spills and the size of every loop body.  usage: scripts/dev_sass.sh <tag> [tier] [extra -D...]
tag=$1; tier=${2:-192}; shift; shift
cd "$(dirname "$0")/../kamphir_b200/csrc"
PK_OUT=/tmp/lib_dev_$tag.so PK_PTXAS_V=1 PK_DEFS="-DPK_DEV_TIER=$tier $*" bash build.sh 2>&1 | grep -A1 "pk_dp_kernel" | grep -v "^--" | grep "registers\|spill" 
cuobjdump -sass /tmp/lib_dev_$tag.so | awk '/Function : .*pk_dp_kernel/{f=1} /Function : .*pk_zero/{f=0} f' > /tmp/dev_$tag.sass
python3 - /tmp/dev_$tag.sass <<'PY'
import re, sys
ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: k for k, (a, _) in enumerate(ins)}
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:`\()?\.?L?_?x?_?\w*\)?\s*(0x[0-9a-f]+)?", t)
    m2 = re.search(r"BRA.*?0x([0-9a-f]+)", t)
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt <= a and tgt in addr:
            body = ins[addr[tgt]:k + 1]
            n = len(body)
            if n >= 20:
                d = sum(1 for _, x in body if re.match(r"(@!?U?P\d\s+)?D(ADD|MUL|FMA)", x))
                l = sum(1 for _, x in body if re.search(r"\bLDG", x))
                s = sum(1 for _, x in body if re.search(r"\bSTG", x))
                print("loop %05x-%05x  %3d instr  fp64 %d  LDG %d  STG %d  LDS %d  LDC %d" % (tgt, a, n, d, l, s,
                      sum(1 for _, x in body if "LDS" in x), sum(1 for _, x in body if "LDC" in x)))
PY
