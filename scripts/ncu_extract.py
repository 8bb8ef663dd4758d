"""Turns one `ncu --set full` capture of the DP kernel into the two files under profiles/ that bench.py and the summary
cite:  python scripts/ncu_extract.py gpurun_out/prof.ncu-rep <dps in the captured launch> <label> <prefix> <tips> <precision>
writes profiles/<prefix>_ncu_full_selected.csv, profiles/<prefix>_dram_traffic.json and (when the report carries source
counters) profiles/<prefix>_stall_lines.csv  (developer tool; needs the `ncu` CLI, no GPU)."""
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__average_warps_issue_stalled_")


def page(rep, which):
    """CSV of one page of a report: from the .ncu-rep itself, or from <rep>.<page>.csv exported on the GPU box (the reports
    are ~23 MB each with sources imported; scripts/profile_r02.sh exports the two pages there and drops the report)."""
    if rep.endswith(".ncu-rep"):
        return subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return open("%s.%s.csv" % (rep, which)).read()


def stall_lines(rep, prefix, label, top=40):
    """The SASS lines with the most warp-stall samples (source page of a capture taken with --import-source on)."""
    raw = page(rep, "source")
    rows = list(csv.reader(io.StringIO(raw)))
    while rows and not (len(rows[0]) > 2 and rows[0][0].strip().lower() == "address"):      # "Kernel Name", ... line(s) first
        rows.pop(0)
    if len(rows) < 3:
        return 0
    hdr = rows[0]

    def col(name):
        for k, h in enumerate(hdr):
            if h.strip().lower() == name:
                return k
        return -1
    c_src, c_smp, c_exe, c_adr = col("source"), col("warp stall sampling (all samples)"), col("instructions executed"), col("address")
    if c_src < 0 or c_smp < 0:
        return 0
    body = []
    for r in rows[1:]:
        try:
            body.append((float(r[c_smp] or 0), r))
        except (ValueError, IndexError):
            continue
    total = sum(b[0] for b in body) or 1.0
    insts = sum(float(r[c_exe] or 0) for _, r in body if c_exe >= 0 and r[c_exe].replace(".", "").isdigit())
    body.sort(key=lambda b: -b[0])
    with open("profiles/%s_stall_lines.csv" % prefix, "w", newline="") as f:
        f.write('"# ncu --set full --import-source on, source page (%s): the %d SASS lines with the most warp-stall samples; '
                'total samples %d, warp instructions %d"\n' % (label.replace('"', "'"), top, total, insts))
        w = csv.writer(f)
        w.writerow(("offset", "samples_pct", "warp_instructions_executed", "sass"))
        for smp, r in body[:top]:
            w.writerow((r[c_adr] if c_adr >= 0 else "", "%.2f" % (100.0 * smp / total), r[c_exe] if c_exe >= 0 else "", r[c_src]))
    return len(body)


def main():
    rep, dps, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    prefix = sys.argv[4]
    tips, prec = int(sys.argv[5]), int(sys.argv[6])
    raw = page(rep, "raw")
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [("metric", "unit", "value")]
    got = {}
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or (k.endswith("_") and h.startswith(k) and "not_issued" not in h.lower()) for k in KEEP):
            out.append((h, u, v))
            got[h] = (u, v)
    with open("profiles/%s_ncu_full_selected.csv" % prefix, "w", newline="") as f:
        csv.writer(f).writerows(out)

    def to_bytes(name):
        u, v = got[name]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    u, v = got["gpu__time_duration.sum"]
    ms = float(v.replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u]
    json.dump({"command": label, "dps_in_launch": dps, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "kernel_ms_under_ncu": ms, "n_tips": tips, "precision": prec, "dram_bytes_per_dp": (rd + wr) / dps},
              open("profiles/%s_dram_traffic.json" % prefix, "w"), indent=1)
    n = stall_lines(rep, prefix, label)
    print("wrote profiles/%s_ncu_full_selected.csv (%d metrics), profiles/%s_dram_traffic.json, stall lines from %d SASS lines"
          % (prefix, len(out) - 1, prefix, n))


if __name__ == "__main__":
    main()
