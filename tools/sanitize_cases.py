#!/usr/bin/env python
"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck) on a B200: every kernel family
runs at least once -- two-stage and generic temporal kernels, TMA reduce, band kernel with and without blur, with
the heat map (raw and coloured) and in feature mode, baseband, finalize, pooling, host streaming path, the float /
16-bit and planar-YUV front ends, file descriptors, resize."""
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
# float / 16-bit front-end bodies of the two-stage kernel (fp32 sRGB, fp16 linear, uint16 PQ)
tst, ref = synth.make_pair_u8(11, 5, 48, 128)
m = cv.cvvdp(display_name="standard_fhd", device=dev)
out.append(float(m.predict(t(tst).float() / 255, t(ref).float() / 255, frames_per_second=30)[0]))
m = cv.cvvdp(display_name="standard_hdr_linear", device=dev)
out.append(float(m.predict((t(tst).float() * 2).half(), (t(ref).float() * 2).half(), frames_per_second=30)[0]))
m = cv.cvvdp(display_name="standard_hdr_pq", device=dev)
out.append(float(m.predict((t(tst).to(torch.int32) * 200).to(torch.int16), (t(ref).to(torch.int32) * 200).to(torch.int16), frames_per_second=30)[0]))
# 8-bit image batches: the table front-end kernel, planar and interleaved; a gray image
tst, ref = synth.make_pair_u8(14, 1, 64, 128)
m = cv.cvvdp(display_name="standard_fhd", device=dev)
out.append(float(m.predict(t(tst)[:, :, 0], t(ref)[:, :, 0], dim_order="BCHW")[0]))
out.append(float(m.predict(t(tst)[0, :, 0].permute(1, 2, 0).contiguous(), t(ref)[0, :, 0].permute(1, 2, 0).contiguous(), dim_order="HWC")[0]))
out.append(float(m.predict(t(tst)[0, 0, 0], t(ref)[0, 0, 0], dim_order="HW")[0]))
# long filters (120 and 90 fps: 31 and 25 taps): shared-memory-ring temporal kernel, table and float variants
tst, ref = synth.make_pair_u8(13, 37, 32, 128)
m = cv.cvvdp(display_name="standard_fhd", device=dev, temp_padding="symmetric")
out.append(float(m.predict(t(tst), t(ref), frames_per_second=120)[0]))
out.append(float(m.predict(t(tst).float() / 255, t(ref).float() / 255, frames_per_second=90)[0]))
# planar YUV: cp.async-staged rows of the two-stage kernel (4:2:0 8 bit, 4:2:2 10 bit, 4:4:4), file descriptors, windows;
# another width through the generic kernel; full-screen resize (all four filters) frame by frame
import tempfile  # noqa: E402
from golden.make_golden_yuv_synth import synth_yuv  # noqa: E402
with tempfile.TemporaryDirectory() as td:
    for chroma, bd, cs, disp, W in (("420", 8, "709", "standard_fhd", 128), ("422", 10, "2020", "standard_hdr_pq", 192),
                                    ("444", 8, "709", "standard_fhd", 64), ("420", 8, "709", "standard_fhd", 100)):
        F, H = 7, 36
        ty, ry = synth_yuv(12, F, H, W, chroma, bd)
        props = {"width": W, "height": H, "fps": 30, "bit_depth": bd, "color_space": cs, "chroma_ss": chroma}
        tf, rf = os.path.join(td, cv.create_yuv_fname("t" + chroma + str(W), props)), os.path.join(td, cv.create_yuv_fname("r" + chroma + str(W), props))
        ty.tofile(tf), ry.tofile(rf)
        m = cv.cvvdp(display_name=disp, device=dev)
        m.yuv_chunk_bytes = 1
        out.append(float(m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry=disp))[0]))
    for mode in ("nearest", "bilinear", "bicubic", "area"):
        vs = cv.video_source_yuv_file(tf, rf, display_photometry="standard_fhd", frames=2, full_screen_resize=mode, resize_resolution=(150, 77))
        out.append(float(m.predict_video_source(vs)[0]))
torch.cuda.synchronize()
print("sanitize cases OK:", ["%.5f" % v for v in out])
