#!/bin/bash
# Same-box A/B of library builds under build/ab/*.so: u8 and f32 headline shape, temporal kernel and step times.
cp colorvideovdp_b200/libcvvdp_b200.so /tmp/lib_keep.so
for rep in 1; do
for lib in build/ab/*.so; do
  cp $lib colorvideovdp_b200/libcvvdp_b200.so
  for dt in u8 f32; do
    python bench.py --dtype $dt --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > /tmp/ab.json
    python - "$lib" "$dt" <<'PY'
import json, sys
d = json.loads(open('/tmp/ab.json').read())
k = d['kernels']
print(f"{sys.argv[1]:28s} {sys.argv[2]:4s} step {d['ms_per_step']:7.3f} ms  temporal {k['temporal']['ms_per_step']:6.3f}  band_l0 {k['band_l0']['ms_per_step']:6.3f}  reduce_l0 {k['reduce_l0']['ms_per_step']:6.3f}")
PY
  done
done
done
cp /tmp/lib_keep.so colorvideovdp_b200/libcvvdp_b200.so
