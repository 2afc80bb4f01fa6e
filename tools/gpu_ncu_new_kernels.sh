#!/bin/bash
# ncu --set full digests of the temporal-kernel variants added late in round 2 (one launch each)
mkdir -p gpurun_out
cat > /tmp/hfr.py <<'PY'
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import colorvideovdp_b200 as cv, synth
dev = torch.device('cuda:0')
t, r = synth.make_pair_u8(5, 8, 1080, 1920)
td = torch.from_numpy(t).to(dev).repeat(1, 1, 8, 1, 1)[:, :, :60].contiguous(); rd = torch.from_numpy(r).to(dev).repeat(1, 1, 8, 1, 1)[:, :, :60].contiguous()
m = cv.cvvdp(display_name='standard_fhd', device=dev)
print(float(m.predict(td, rd, frames_per_second=120)[0]))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_temporal_sr -c 1 -f -o /tmp/p_sr python /tmp/hfr.py > gpurun_out/ncu_sr.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_temporal_2s -c 1 -f -o /tmp/p_f32 python bench.py --dtype f32 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_f32.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_temporal_2s -c 1 -f -o /tmp/p_yuv python tools/probe_yuv_e2e.py > gpurun_out/ncu_yuv.txt 2>&1
for k in sr f32 yuv; do
  ncu -i /tmp/p_$k.ncu-rep --page raw --csv > gpurun_out/raw_t_$k.csv 2>/dev/null
  ncu -i /tmp/p_$k.ncu-rep --page source --csv > gpurun_out/src_t_$k.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
