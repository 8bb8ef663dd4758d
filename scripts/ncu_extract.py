"""Turns one `ncu --set full` capture of the DP kernel into the two files under profiles/ that bench.py and the summary
cite:  python scripts/ncu_extract.py gpurun_out/prof.ncu-rep <dps in the captured launch> <label> [file suffix]
(developer tool; needs the `ncu` CLI, no GPU)."""
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


def main():
    rep, dps, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    suffix = sys.argv[4] if len(sys.argv) > 4 else ""          # e.g. "_fp32": a second capture next to the fp64 one
    prec = 32 if "32" in suffix else 64
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [("metric", "unit", "value")]
    got = {}
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or (k.endswith("_") and h.startswith(k) and "not_issued" not in h.lower()) for k in KEEP):
            out.append((h, u, v))
            got[h] = (u, v)
    with open("profiles/r01_dp_kernel%s_ncu_full_selected.csv" % suffix, "w", newline="") as f:
        csv.writer(f).writerows(out)

    def to_bytes(name):
        u, v = got[name]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    u, v = got["gpu__time_duration.sum"]
    ms = float(v) * {"ms": 1, "us": 1e-3, "s": 1e3}[u]
    json.dump({"command": label, "dps_in_launch": dps, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "kernel_ms_under_ncu": ms, "n_tips": 500, "precision": prec, "dram_bytes_per_dp": (rd + wr) / dps},
              open("profiles/r01_dram_traffic%s.json" % suffix, "w"), indent=1)
    print("wrote profiles/r01_dp_kernel%s_ncu_full_selected.csv (%d metrics) and profiles/r01_dram_traffic%s.json" % (suffix, len(out) - 1, suffix))


if __name__ == "__main__":
    main()
