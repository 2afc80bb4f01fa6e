#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck) on a B200: every kernel family
runs at least once -- two-stage and generic temporal kernels, TMA reduce, band kernel with and without blur, with
the heat map (raw and coloured) and in feature mode, baseband, finalize, pooling, host streaming path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import colorvideovdp_b200 as cv  # noqa: E402
import synth  # noqa: E402

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
out = []
tst, ref = synth.make_pair_u8(7, 6, 72, 128)          # two-stage temporal kernel, 9 taps
m = cv.cvvdp(display_name="standard_fhd", device=dev)
out.append(float(m.predict(t(tst), t(ref), frames_per_second=30)[0]))
out.append(float(m.predict(tst, ref, frames_per_second=30)[0]))  # host streaming path (bounce buffers: pageable numpy)
tst, ref = synth.make_pair_u8(8, 19, 130, 250)         # 17 taps, ragged strips, generic temporal kernel (npix % 64 != 0)
m = cv.cvvdp(display_name="standard_4k", device=dev, heatmap="raw")
out.append(float(m.predict(t(tst), t(ref), frames_per_second=60)[0]))
m = cv.cvvdp(display_name="standard_4k", device=dev, heatmap="supra-threshold")
out.append(float(m.predict(t(tst[:, :, :4]), t(ref[:, :, :4]), frames_per_second=60)[0]))
tst, ref = synth.make_pair_u8(9, 2, 96, 160)
m = cv.cvvdp(display_name="standard_fhd", device=dev)
vs = cv.video_source_array(t(tst), t(ref), 30, display_photometry=m.display_photometry)
feats, _ = m.extract_features(vs)
out.append(float(feats[0].sum()))
img_t, img_r = synth.make_pair_u8(10, 1, 40, 56)
out.append(float(m.predict(t(img_t), t(img_r))[0]))   # image: three channels, tiny levels without blur
torch.cuda.synchronize()
print("sanitize cases OK:", ["%.5f" % v for v in out])
