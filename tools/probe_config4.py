#!/usr/bin/env python
"""Where does the time go in config 4 (4K HDR PQ, raw heat map) through predict() on pinned host tensors?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import colorvideovdp_b200 as cv
from colorvideovdp_b200 import cvvdp_metric as CM
dev = torch.device("cuda:0")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 60
m = cv.cvvdp(display_name="standard_hdr_pq", device=dev, heatmap="raw")
tst, ref = bench.make_clip(3, 0, F, 2160, 3840, "u16", dev, hdr=True)
th, rh = bench.pinned_like(tst), bench.pinned_like(ref)
del tst, ref
orig = m._ctx.process_host
def timed_ph(*a):
    t0 = time.perf_counter(); r = orig(*a); torch.cuda.synchronize(); print(f"  process_host {1e3*(time.perf_counter()-t0):.1f} ms"); return r
m._ctx.process_host = timed_ph
for i in range(3):
    t0 = time.perf_counter()
    jod, stats = m.predict(th, rh, dim_order="BCFHW", frames_per_second=60.0)
    j = float(jod)
    print(f"predict {1e3*(time.perf_counter()-t0):.1f} ms  jod {j:.5f} pinned heatmap: {stats['heatmap'].is_pinned()}")
    del stats
