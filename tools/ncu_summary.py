#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into markdown for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_%"),
    ("smsp__issue_active.avg.per_cycle_active", "issue_active"),
    ("smsp__inst_executed.sum", "warp_instr"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("ratio")]
    print(f"# ncu --set full summary of `{rep}`\n")
    print("Durations are cold-cache, serialised profiler replays: compare shares, not absolutes.\n")
    for r in data:
        name = r[ix["Kernel Name"]]
        print(f"## {name}\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in ix:
                print(f"| {label} | {r[ix[k]]} {units[ix[k]]} |")
        vals = []
        for h in stall:
            try:
                vals.append((float(r[ix[h]]), h))
            except ValueError:
                pass
        top = sorted(vals, reverse=True)[:6]
        txt = ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
                                     .replace("_per_issue_active.ratio", ""), v) for v, h in top)
        print(f"| top stalls (cycles per issued instruction) | {txt} |\n")


if __name__ == "__main__":
    main()
