#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_<name>.csv

One row per profiled launch with the metrics the roofline discussion uses (duration, DRAM
bytes read/written, DRAM / L2 / L1 / SM throughput as % of peak, occupancy, registers)."""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [c for c in WANT if c in hdr]
    idx = [hdr.index(c) for c in cols]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{c} [{units[i]}]" if units[i] else c for c, i in zip(cols, idx)])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(f"{len(rows) - 2} launches -> {out}")


if __name__ == "__main__":
    main()
