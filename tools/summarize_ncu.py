#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  launches:  tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/<tag>_launches.csv
             per-kernel launch count, total / mean device time and share of all cvvdp kernels
             (from `ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
  kernel:    tools/summarize_ncu.py kernel gpurun_out/prof_k_band.ncu-rep profiles/<tag>_k_band.csv
             the metrics the roofline discussion uses, from one `ncu --set full` capture.
"""
import csv
import re
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__cycles_elapsed.max",
]


def launches(src, dst):
    agg = {}
    with open(src, newline="") as fh:
        rows = [r for r in csv.reader(fh) if len(r) >= 15 and r[0].isdigit()]
    for r in rows:
        name, metric, unit, val = r[4], r[12], r[13], r[14]
        if metric != "gpu__time_duration.sum":
            continue
        ms = float(val.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        short = re.sub(r"\(.*", "", name)
        ours = short.startswith("cvvdp::") or re.match(r"(void )?k_[a-z0-9_]+", short)
        if not ours:
            short = "other: " + short[:60]
        else:
            short = "cvvdp::" + re.sub(r"^(void )?(cvvdp::)?", "", short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ms
    ours = sum(v[1] for k, v in agg.items() if k.startswith("cvvdp::")) or 1.0
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches", "total_ms", "mean_ms", "share_of_cvvdp_kernels"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f"{v[1]:.4f}", f"{v[1] / v[0]:.4f}", f"{v[1] / ours:.4f}" if k.startswith("cvvdp::") else ""])


def kernel(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            name = re.sub(r"\(.*", "", vals[hdr.index("Kernel Name")])
            for m in KEY_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    w.writerow([name, m, units[i], vals[i]])


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
