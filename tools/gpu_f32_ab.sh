#!/bin/bash
# A/B for the float-input front end of the temporal kernel: parity tests, then the f32 headline shape and the extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 > gpurun_out/pytest_parity.txt 2>&1; tail -3 gpurun_out/pytest_parity.txt
python bench.py --dtype f32 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_f32.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_f32.txt').read())
print('f32', d['value'], d['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>1})
PY
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_extras.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_extras.txt').read())
print('u8', d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d.get('extra',{}).items(): print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','temporal_kernel','e2e','unavailable')})
PY
