"""Golden fixtures for raw planar YUV ingestion, generated with the UNMODIFIED reference
(pycvvdp.video_source_yuv.video_source_yuv_file + cvvdp.predict_video_source, CPU).  Container only:

    python tests/golden/make_golden_yuv.py

Each tests/golden/yuv_*.npz stores the raw bytes of the synthetic test/reference .yuv files, their file
names (the metadata lives in the name), the display, and the reference outputs (JOD, Q_per_ch, and the
RGB tensor of one frame from YUVReader.get_frame_rgb_tensor)."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402
import synth  # noqa: E402

pycvvdp = ref_loader.load()
from pycvvdp import video_source_yuv as ref_yuv  # noqa: E402

DEV = torch.device("cpu")


from make_golden_yuv_synth import synth_yuv  # noqa: E402


def save(name, seed, F, H, W, fps, chroma, bit_depth, color_space, display, padding="replicate"):
    t, r = synth_yuv(seed, F, H, W, chroma, bit_depth)
    props = {"width": W, "height": H, "fps": fps, "bit_depth": bit_depth, "color_space": color_space, "chroma_ss": chroma}
    with tempfile.TemporaryDirectory() as td:
        tf = os.path.join(td, ref_yuv.create_yuv_fname("test", props))
        rf = os.path.join(td, ref_yuv.create_yuv_fname("ref", props))
        t.tofile(tf)
        r.tofile(rf)
        vs = ref_yuv.video_source_yuv_file(tf, rf, display_photometry=display)
        m = pycvvdp.cvvdp(display_name=display, device=DEV, temp_padding=padding, quiet=True)
        with torch.no_grad():
            q, s = m.predict_video_source(vs)
            rgb = vs.test_vidr.get_frame_rgb_tensor(F - 1, DEV).numpy()
        meta = {"test_name": os.path.basename(tf), "ref_name": os.path.basename(rf), "display": display, "padding": padding}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), test_bytes=t, ref_bytes=r, meta=np.asarray(json.dumps(meta)),
                        jod=np.asarray(q.numpy(), dtype=np.float32), Q_per_ch=s["Q_per_ch"].astype(np.float32),
                        rgb_last_test_frame=rgb.astype(np.float32))
    print(name, float(q), s["Q_per_ch"].shape)


if __name__ == "__main__":
    save("yuv_420_8b_709_6x48x64", 61, 6, 48, 64, 24, "420", 8, "709", "standard_fhd")
    save("yuv_444_10b_2020_4x36x52_hdr", 62, 4, 36, 52, 30, "444", 10, "2020", "standard_hdr_pq")
    save("yuv_422_8b_709_12x40x48_sym", 63, 12, 40, 48, 25, "422", 8, "709", "standard_4k", "symmetric")
    save("yuv_420_10b_2020_1x64x96_image", 64, 1, 64, 96, 30, "420", 10, "2020", "standard_hdr_pq")
