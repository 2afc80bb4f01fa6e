#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): the evidence that the kernels use the
Blackwell paths they claim -- UTMALDG (cp.async.bulk.tensor), UBLKCP, LDGSTS (cp.async), FFMA2 / FMUL2 / FADD2
(packed fp32), MUFU, SYNCS (mbarrier).  usage: tools/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "colorvideovdp_b200", "libcvvdp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "ATOMG", "RED", "ATOMS"]
total = collections.Counter()
rows = []
for fn in re.split(r"\n\s+Function : ", txt)[1:]:
    name = fn.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void cvvdp::", "").replace("cvvdp::", "")
    ops = collections.Counter(re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", fn))
    n = sum(ops.values())
    rows.append((dem, n, ops))
    total.update(ops)
print(f"# {os.path.basename(lib)}: {len(rows)} kernels, {sum(total.values())} SASS instructions")
print("kernel".ljust(64) + " instrs " + " ".join(k.rjust(7) for k in KEY))
for dem, n, ops in sorted(rows, key=lambda r: r[0]):
    print(dem[:63].ljust(64) + f"{n:7d} " + " ".join(str(ops.get(k, 0)).rjust(7) for k in KEY))
print("TOTAL".ljust(64) + f"{sum(total.values()):7d} " + " ".join(str(total.get(k, 0)).rjust(7) for k in KEY))
