"""Developer tool: per basic block (run of SASS lines with the same execution count) of an exported ncu source page: share of the
executed warp instructions and of the warp-stall samples, plus the kernel's stall reasons per issue from the raw page.
usage: python scripts/ncu_blocks.py gpurun_out/prof_<name>   (reads <prefix>.source.csv and <prefix>.raw.csv)"""
import csv
import sys

pre = sys.argv[1]
rows = list(csv.reader(open(pre + ".source.csv")))
while rows and not (len(rows[0]) > 2 and rows[0][0].strip().lower() == "address"):
    rows.pop(0)
hdr, body = rows[0], rows[1:]
ci = {h: i for i, h in enumerate(hdr)}
A = []
for r in body:
    try:
        A.append((int(r[ci["Address"]], 16), r[ci["Source"]].strip(), float(r[ci["Warp Stall Sampling (All Samples)"]] or 0),
                  int(r[ci["Instructions Executed"]] or 0), float(r[ci["Avg. Threads Executed"]] or 0)))
    except (ValueError, IndexError):
        pass
base = A[0][0]
ts, tn = sum(x[2] for x in A), sum(x[3] for x in A)
print("lines %d, stall samples %d, warp instructions %d" % (len(A), ts, tn))
seg, cur = [], None
for a, src, s, n, thr in A:
    if cur and n == cur["n"]:
        cur["len"] += 1; cur["s"] += s; cur["end"] = a
        if s > cur["top"][0]: cur["top"] = (s, src)
    else:
        if cur: seg.append(cur)
        cur = dict(start=a, end=a, n=n, len=1, s=s, thr=thr, top=(s, src))
seg.append(cur)
for x in seg:
    if x["n"] * x["len"] > tn * 0.004 or x["s"] > ts * 0.004:
        print("%05x-%05x len %3d exec %9d  instr %5.2f%%  stall %5.2f%%  thr %.1f   top: %.2f%% %s" % (
            x["start"] - base, x["end"] - base, x["len"], x["n"], 100 * x["n"] * x["len"] / tn, 100 * x["s"] / ts, x["thr"],
            100 * x["top"][0] / ts, x["top"][1][:44]))
raw = list(csv.reader(open(pre + ".raw.csv")))
h, v = raw[0], raw[2]
for k, val in zip(h, v):
    if ("issue_stalled" in k and "per_issue_active" in k) or k in (
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
            "sm__warps_active.avg.per_cycle_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "launch__registers_per_thread"):
        try:
            if float(val) >= 0.3: print("  %-80s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), val))
        except ValueError:
            pass
