#!/usr/bin/env python
"""Digest the CSV exports of one ncu capture (made on the GPU box with
`ncu -i rep --page raw --csv` and `--page source --csv`): key metrics + SASS opcode mix per barrier phase."""
import collections
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size']


def main(raw, src, units_per_launch=None, out=None):
    o = open(out, "w") if out else sys.stdout
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    print("kernel:", vals[hdr.index("Kernel Name")][:80], file=o)
    for k in KEYS + [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]:
        if k in hdr:
            print(f"{k},{vals[hdr.index(k)]},{units[hdr.index(k)]}", file=o)
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[iI]) for r in data)
    totS = sum(int(r[iN]) for r in data) or 1
    norm = (units_per_launch / 32.0) if units_per_launch else None
    phase, agg, ops = 0, collections.defaultdict(lambda: [0, 0]), collections.defaultdict(lambda: collections.defaultdict(int))
    allops = collections.defaultdict(int)
    for r in data:
        t = r[iS].split()
        op = t[1] if t and t[0].startswith('@') else (t[0] if t else '?')
        op = op.rstrip(';').split('.')[0]
        agg[phase][0] += int(r[iI]); agg[phase][1] += int(r[iN]); ops[phase][op] += int(r[iI]); allops[op] += int(r[iI])
        if op == 'BAR':
            phase += 1
    print(f"total warp-instructions {tot}" + (f" = {tot / norm:.0f} per unit" if norm else ""), file=o)
    fmt = (lambda v: round(v / norm, 1)) if norm else (lambda v: v)
    print("opcode mix:", [(k, fmt(v)) for k, v in sorted(allops.items(), key=lambda kv: -kv[1])[:16]], file=o)
    for p, a in sorted(agg.items()):
        top = sorted(ops[p].items(), key=lambda kv: -kv[1])[:8]
        print(f"phase {p}: inst {100 * a[0] / tot:.1f}% samples {100 * a[1] / totS:.1f}%", [(k, fmt(v)) for k, v in top], file=o)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None, sys.argv[4] if len(sys.argv) > 4 else None)
