#!/bin/bash
# Is the float-input temporal kernel slower inside the default bench (after the u8 runs, in the extras) than standalone?
for i in 1 2; do
python bench.py --dtype f32 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > /tmp/a.json
python - <<'PY'
import json
d = json.loads(open('/tmp/a.json').read()); k = d['kernels']
print(f"standalone f32: step {d['ms_per_step']:.2f} temporal {k['temporal']['ms_per_step']:.2f} band_l0 {k['band_l0']['ms_per_step']:.2f} clocks {d['clocks']}")
PY
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/b.json
python - <<'PY'
import json
d = json.loads(open('/tmp/b.json').read()); k = d['kernels']
print(f"default u8: step {d['ms_per_step']:.2f} temporal {k['temporal']['ms_per_step']:.2f} band_l0 {k['band_l0']['ms_per_step']:.2f} clocks {d['clocks']}")
print("   extras f32:", d['extra']['config3_f32_input']['ms_per_step'], d['extra']['config3_f32_input']['temporal_kernel'])
PY
done
nvidia-smi --query-gpu=name,power.limit,power.draw,clocks.sm,clocks.max.sm,temperature.gpu --format=csv
