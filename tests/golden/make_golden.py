"""Generate tests/golden/*.npz by running the UNMODIFIED reference (pycvvdp imported from
/root/reference, CPU, fp32) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Each .npz holds the exact inputs, the call arguments (JSON string `meta`) and the reference
outputs (`jod`, `Q_per_ch`, `rho_band`, optional `heatmap`, optional stage-level tensors), so the
fixtures do not depend on any random-number generator at test time.  The fixtures pin both the
numpy oracle (tests/test_oracle_golden.py, no GPU) and the CUDA path (tests/test_gpu_parity.py).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402
import synth  # noqa: E402

pycvvdp = ref_loader.load()
DEV = torch.device("cpu")
torch.set_num_threads(8)


def run_ref(test, ref, dim_order, fps, display, padding="replicate", heatmap=None):
    m = pycvvdp.cvvdp(display_name=display, device=DEV, temp_padding=padding, heatmap=heatmap, quiet=True)
    with torch.no_grad():
        q, s = m.predict(test, ref, dim_order=dim_order, frames_per_second=fps)
    out = {"jod": np.asarray(q.cpu().numpy(), dtype=np.float32), "Q_per_ch": s["Q_per_ch"].astype(np.float32),
           "rho_band": np.asarray(s["rho_band"], dtype=np.float64)}
    if heatmap is not None:
        out["heatmap"] = s["heatmap"].numpy()
    return out, m


def save(name, test, ref, dim_order, fps, display, padding="replicate", heatmap=None, extra=None):
    out, _ = run_ref(test, ref, dim_order, fps, display, padding, heatmap)
    meta = {"dim_order": dim_order, "fps": fps, "display": display, "padding": padding, "heatmap": heatmap,
            "reference": "gfxdisp/ColorVideoVDP pycvvdp 0.5.4 (params 0.5.6), torch %s CPU" % torch.__version__}
    arrays = dict(test=test, ref=ref, meta=np.asarray(json.dumps(meta)), **out)
    if extra:
        arrays.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(f"{name}: JOD={out['jod']}  Q_per_ch{out['Q_per_ch'].shape}")


def stage_dump(test, ref, dim_order, fps, display, frame):
    """Stage-level tensors from the reference's own functions for one frame."""
    m = pycvvdp.cvvdp(display_name=display, device=DEV, quiet=True)
    vs = pycvvdp.video_source.video_source_array(test, ref, fps, dim_order=dim_order, display_photometry=m.display_photometry)
    H, W, F = vs.get_video_size()
    m.lpyr = pycvvdp.lpyr_dec.weber_contrast_pyr(W, H, m.pix_per_deg, DEV, contrast=m.contrast)
    ex = {}
    with torch.no_grad():
        ex["st_dkl_test"] = vs.get_test_frame(frame, DEV, "DKLd65").numpy()
        ex["st_dkl_ref"] = vs.get_reference_frame(frame, DEV, "DKLd65").numpy()
        if F > 1:
            m.F, _ = m.get_temporal_filters(fps)
            m.filter_len = torch.numel(m.F[0])
            fb = pycvvdp.cvvdp_metric.cvvdp_frame_buffers()
            R = None
            for ff in range(frame + 1):
                R = m.read_block_of_frames(vs, 4, fb, 1, "DKLd65", ff, 1)
            ex["st_filters"] = torch.stack(m.F).numpy()
        else:
            fb = pycvvdp.cvvdp_metric.cvvdp_frame_buffers()
            R = m.read_block_of_frames(vs, 3, fb, 1, "DKLd65", 0, 1)
        ex["st_R"] = R.numpy()
        gpyr = m.lpyr.gaussian_pyramid_dec(R, m.lpyr.height + 1)
        bands, logL = m.lpyr.decompose(R)
        for i, g in enumerate(gpyr):
            ex[f"st_gpyr{i}"] = g.numpy()
        for i, (b, l) in enumerate(zip(bands, logL)):
            ex[f"st_band{i}"] = m.lpyr.get_band(bands, i).numpy()
            ex[f"st_logL{i}"] = l.numpy()
    ex["st_frame"] = np.asarray(frame)
    return ex


def main():
    # 1. images, odd/even sizes (lpyr_dec.py:206 parity quirk both ways)
    t, r = synth.make_pair_u8(11, 1, 135, 240)
    t, r = t[0, :, 0].transpose(1, 2, 0).copy(), r[0, :, 0].transpose(1, 2, 0).copy()
    save("img_u8_135x240_fhd", t, r, "HWC", 0, "standard_fhd")
    t, r = synth.make_pair_u8(12, 1, 70, 121)
    t, r = t[0, :, 0].transpose(1, 2, 0).copy(), r[0, :, 0].transpose(1, 2, 0).copy()
    save("img_u8_70x121_4k", t, r, "HWC", 0, "standard_4k",
         extra=stage_dump(t, r, "HWC", 0, "standard_4k", 0))
    # 2. videos: replicate / symmetric padding, 30 fps (fl=9)
    t, r = synth.make_pair_u8(13, 10, 64, 100)
    save("vid_u8_10x64x100_fhd_rep", t, r, "BCFHW", 30, "standard_fhd", "replicate",
         extra=stage_dump(t, r, "BCFHW", 30, "standard_fhd", 9))
    save("vid_u8_10x64x100_fhd_sym", t, r, "BCFHW", 30, "standard_fhd", "symmetric")
    # 3. HDR PQ uint16, 60 fps (fl=17)
    t, r = synth.make_pair_pq_u16(14, 18, 48, 80)
    save("vid_u16_18x48x80_hdrpq_60", t, r, "BCFHW", 60, "standard_hdr_pq")
    # 4. clip shorter than the filter, symmetric ping-pong
    t, r = synth.make_pair_u8(15, 5, 40, 64)
    save("vid_u8_5x40x64_fhd_sym_short", t, r, "BCFHW", 30, "standard_fhd", "symmetric")
    save("vid_u8_5x40x64_fhd_rep_short", t, r, "BCFHW", 30, "standard_fhd", "replicate")
    # 5. batch of 2 fp32 images against a single reference (singleton batch broadcast)
    t0, r0 = synth.make_pair_u8(16, 1, 40, 56)
    t1, _ = synth.make_pair_u8(17, 1, 40, 56, noise_sigma=9.0)
    tb = (np.concatenate([t0[:, :, 0], t1[:, :, 0]], 0).astype(np.float32) / 255).astype(np.float32)
    rb = (r0[:, :, 0].astype(np.float32) / 255).astype(np.float32)
    save("img_f32_b2_40x56_4k", tb, rb, "BCHW", 0, "standard_4k")
    # 6. grey-scale fp16 image "HW" on a linear HDR display
    t, r = synth.make_pair_u8(18, 1, 50, 70, C=1)
    tl = (t[0, 0, 0].astype(np.float32) * 4.0).astype(np.float16)
    rl = (r[0, 0, 0].astype(np.float32) * 4.0).astype(np.float16)
    save("img_f16_gray_50x70_hdrlin", tl, rl, "HW", 0, "standard_hdr_linear")
    # 7. raw heat maps (partition independent)
    t, r = synth.make_pair_u8(19, 1, 64, 96)
    save("img_u8_64x96_4k_hmraw", t, r, "BCFHW", 0, "standard_4k", heatmap="raw")
    t, r = synth.make_pair_u8(20, 8, 48, 64)
    save("vid_u8_8x48x64_fhd_24_hmraw", t, r, "BCFHW", 24, "standard_fhd", heatmap="raw")
    # 7b. coloured heat maps of an image (one block in the reference => partition independent)
    t, r = synth.make_pair_u8(22, 1, 48, 72)
    save("img_u8_48x72_4k_hmthr", t, r, "BCFHW", 0, "standard_4k", heatmap="threshold")
    save("img_u8_48x72_4k_hmsupra", t, r, "BCFHW", 0, "standard_4k", heatmap="supra-threshold")
    # 8. HLG + gamma EOTF coverage (fp32 video, 25 fps -> fl=9)
    t, r = synth.make_pair_u8(21, 3, 36, 52)
    tf, rf = (t.astype(np.float32) / 255), (r.astype(np.float32) / 255)
    save("vid_f32_3x36x52_hlg_25", tf, rf, "BCFHW", 25, "standard_hdr_hlg")
    # 9. crop of the reference's example image with its example distortion (ex_simple_image.py:33-34)
    import cv2
    from scipy.ndimage import gaussian_filter
    im = cv2.imread("/root/reference/example_media/wavy_facade.png", cv2.IMREAD_UNCHANGED)[:, :, ::-1]
    blur = np.zeros_like(im)
    for cc in range(3):
        blur[..., cc] = gaussian_filter(im[..., cc], 2, mode="nearest", truncate=2.0)
    y0, x0 = 256, 384
    save("img_u16_wavy_crop_171x256_4k", blur[y0:y0 + 171, x0:x0 + 256].copy(), im[y0:y0 + 171, x0:x0 + 256].copy(),
         "HWC", 0, "standard_4k")
    # the full known-answer value (docstring 8.514) is asserted container-only in test_oracle_golden.py
    out, _ = run_ref(blur, np.ascontiguousarray(im), "HWC", 0, "standard_4k")
    print("wavy_facade blur full:", out["jod"])
    np.savez_compressed(os.path.join(HERE, "known_answer_wavy_facade.npz"), jod=out["jod"],
                        Q_per_ch=out["Q_per_ch"], docstring_jod=np.asarray(8.514))


if __name__ == "__main__":
    main()
