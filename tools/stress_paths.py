"""Consistency of the execution paths at sizes where the host ring wraps, plans split into several blocks and views are
permuted: every variant of one input must give bit-identical Q_per_ch (the summation order depends on the level
geometry only) and heat maps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import colorvideovdp_b200 as cv  # noqa: E402
import synth  # noqa: E402

dev = torch.device("cuda:0")
bad = []


def clip(F, H, W, seed=5, B=1):
    items = []
    for b in range(B):
        t, r = synth.make_pair_u8(seed + b, min(F, 12), H, W)
        reps = -(-F // t.shape[2])
        items.append((np.tile(t, (1, 1, reps, 1, 1))[:, :, :F], np.tile(r, (1, 1, reps, 1, 1))[:, :, :F]))
    t = np.ascontiguousarray(np.concatenate([i[0] for i in items], 0))
    r = np.ascontiguousarray(np.concatenate([i[1] for i in items], 0))
    # make the frames differ along time (tiling alone repeats them)
    ramp = (np.arange(F, dtype=np.int32) % 7).reshape(1, 1, F, 1, 1)
    t = np.clip(t.astype(np.int32) + ramp, 0, 255).astype(np.uint8)
    return t, r


def check(name, base, other, hm=False):
    same = np.array_equal(base["Q_per_ch"], other["Q_per_ch"])
    if hm:
        same = same and torch.equal(base["heatmap"], other["heatmap"])
    print(f"{'ok ' if same else 'MISMATCH'} {name}")
    if not same:
        d = np.abs(base["Q_per_ch"].astype(np.float64) - other["Q_per_ch"])
        bad.append((name, float(d.max())))


# 1. 1080p, 400 frames @60: device / pinned host / pageable host / small blocks / frame ranges
t, r = clip(400, 1080, 1920)
m = cv.cvvdp(display_name="standard_fhd", device=dev)
_, base = m.predict(torch.from_numpy(t).to(dev), torch.from_numpy(r).to(dev), frames_per_second=60)
_, s = m.predict(torch.from_numpy(t).pin_memory(), torch.from_numpy(r).pin_memory(), frames_per_second=60)
check("1080p x 400 @60: pinned host vs device", base, s)
_, s = m.predict(t, r, frames_per_second=60)
check("1080p x 400 @60: pageable numpy vs device", base, s)
m_small = cv.cvvdp(display_name="standard_fhd", device=dev, gpu_mem=1.0)  # ~1 GB of workspace: many blocks
_, s = m_small.predict(torch.from_numpy(t).to(dev), torch.from_numpy(r).to(dev), frames_per_second=60)
check("1080p x 400 @60: 1 GB workspace (many blocks), device", base, s)
_, s = m_small.predict(t, r, frames_per_second=60)
check("1080p x 400 @60: 1 GB workspace, pageable host", base, s)
vs = cv.video_source_array(t, r, 60, display_photometry=m.display_photometry)
parts = np.zeros_like(base["Q_per_ch"])
for lo, hi in ((0, 37), (37, 38), (38, 251), (251, 400)):
    _, s = m.predict_video_source(vs, frame_range=(lo, hi))
    parts[:, :, lo:hi] = s["Q_per_ch"][:, :, lo:hi]
check("1080p x 400 @60: four frame ranges, host", base, {"Q_per_ch": parts})
# permuted host layout (frames, height, width, channels)
tp, rp = np.ascontiguousarray(t[0].transpose(1, 2, 3, 0)), np.ascontiguousarray(r[0].transpose(1, 2, 3, 0))
_, s = m.predict(tp, rp, dim_order="FHWC", frames_per_second=60)
jod_p = s
_, s2 = m.predict(torch.from_numpy(tp).to(dev), torch.from_numpy(rp).to(dev), dim_order="FHWC", frames_per_second=60)
check("1080p x 400 @60: FHWC numpy vs FHWC device", s2, jod_p)
d = np.abs(s2["Q_per_ch"].astype(np.float64) - base["Q_per_ch"]) / (1e-3 * np.abs(base["Q_per_ch"]) + 1e-5)
print(f"     FHWC (interleaved stage) vs BCFHW (planar stage), both the two-stage kernel: max err/gate {d.max():.4f}")
if d.max() > 1:
    bad.append(("FHWC vs BCFHW", float(d.max())))
del t, r, tp, rp

# 2. 4K HDR, 40 frames, raw heat map: one block vs several, device vs host
t, r = clip(40, 2160, 3840, seed=9)
t16, r16 = (t.astype(np.uint16) * 200).view(np.int16), (r.astype(np.uint16) * 200).view(np.int16)
m = cv.cvvdp(display_name="standard_hdr_pq", device=dev, heatmap="raw")
_, base = m.predict(torch.from_numpy(t16).to(dev), torch.from_numpy(r16).to(dev), frames_per_second=60)
_, s = m.predict(torch.from_numpy(t16).pin_memory(), torch.from_numpy(r16).pin_memory(), frames_per_second=60)
check("4K HDR x 40 raw heat map: pinned host vs device", base, s, hm=True)
m_small = cv.cvvdp(display_name="standard_hdr_pq", device=dev, heatmap="raw", gpu_mem=4.0)
_, s = m_small.predict(torch.from_numpy(t16).to(dev), torch.from_numpy(r16).to(dev), frames_per_second=60)
check("4K HDR x 40 raw heat map: 4 GB workspace (several blocks), device", base, s, hm=True)
_, s = m_small.predict(t16, r16, frames_per_second=60)
check("4K HDR x 40 raw heat map: 4 GB workspace, pageable host", base, s, hm=True)
del t, r, t16, r16, base, s

# 3. batch of 3 x 720p x 90 frames @30, singleton reference batch broadcast, host vs device
t, r = clip(90, 720, 1280, seed=20, B=3)
m = cv.cvvdp(display_name="standard_fhd", device=dev)
_, base = m.predict(torch.from_numpy(t).to(dev), torch.from_numpy(r[:1]).to(dev), frames_per_second=30)
_, s = m.predict(t, r[:1], frames_per_second=30)
check("3 x 720p x 90 @30, reference broadcast: pageable host vs device", base, s)
singles = np.concatenate([m.predict(torch.from_numpy(t[b:b + 1]).to(dev), torch.from_numpy(r[:1]).to(dev), frames_per_second=30)[1]["Q_per_ch"] for b in range(3)], 0)
check("3 x 720p x 90 @30: batch vs one item at a time", base, {"Q_per_ch": singles})

print("stress paths:", "ALL CONSISTENT" if not bad else f"{len(bad)} PROBLEMS {bad}")
