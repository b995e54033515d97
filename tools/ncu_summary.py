#!/usr/bin/env python
"""Text summary of an ncu report (one block per profiled launch): `python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [title]`.
Reads the report through `ncu -i ... --page raw --csv`; the metrics are the ones profiles/README.md quotes."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== ", r[idx["Kernel Name"]].replace("dctb::<", "").replace("(anonymous namespace)::", ""))
        for w in WANT:
            if w in idx:
                print(f"   {w:72s} {r[idx[w]]} {units[idx[w]]}")
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
                stalls.append((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                               float(r[i].replace(",", ""))))
        stalls.sort(key=lambda x: -x[1])
        print("   top stalls (warps per issue): " + ", ".join(f"{h} {v:.2f}" for h, v in stalls[:6]))


if __name__ == "__main__":
    main()
