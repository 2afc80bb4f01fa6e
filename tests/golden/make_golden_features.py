"""Generate tests/golden/feat_*.npz: per-band feature tensors of the UNMODIFIED reference
(pycvvdp.cvvdp_ml_metric.cvvdp_ml_base.extract_features, cvvdp_ml_metric.py:206-298, with
cvvdp_feature_pooling, 78-106), CPU fp32, on small seeded inputs.  Build container only:

    python tests/golden/make_golden_features.py

The ML heads themselves need network weights (hf_hub_download) and are out of scope; the feature
tensors are what the CUDA path has to deliver to them.
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
sys.path.insert(0, os.path.join(ref_loader.REFERENCE_ROOT, "pycvvdp"))  # cvvdp_ml_metric does `from interp import ...`
import pycvvdp.cvvdp_ml_metric as M  # noqa: E402
from pycvvdp.video_source import video_source_array  # noqa: E402

torch.set_num_threads(8)


class FeaturesOnly(M.cvvdp_ml_base):
    """The reference's feature extractor without a regression head."""

    def get_nets_to_load(self):
        return []

    def do_pooling_and_jods(self, features):
        return torch.zeros(1)


def save(name, test, ref, dim_order, fps, display, padding="replicate"):
    m = FeaturesOnly(random_init=True, display_name=display, device=torch.device("cpu"), temp_padding=padding, quiet=True)
    vs = video_source_array(test, ref, fps, dim_order=dim_order, display_photometry=m.display_photometry)
    with torch.no_grad():
        feats, _ = m.extract_features(vs)
    meta = {"dim_order": dim_order, "fps": fps, "display": display, "padding": padding, "n_bands": len(feats),
            "feature_size": int(np.ceil(m.pix_per_deg)),
            "reference": "gfxdisp/ColorVideoVDP pycvvdp 0.5.4 cvvdp_ml_base.extract_features, torch %s CPU" % torch.__version__}
    arrays = {"test": test, "ref": ref, "meta": np.asarray(json.dumps(meta))}
    for bb, f in enumerate(feats):
        arrays[f"features_b{bb}"] = f.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, [tuple(f.shape) for f in feats])


def save_temporal_filters():
    """cvvdp.get_temporal_filters (cvvdp_metric.py:1057-1092) at common frame rates."""
    m = pycvvdp.cvvdp(display_name="standard_4k", device=torch.device("cpu"), quiet=True)
    arrays = {}
    for fps in (15, 24, 25, 30, 50, 59.94, 60, 120):
        F, omega = m.get_temporal_filters(fps)
        arrays[f"fps_{fps}"] = np.stack([f.numpy() for f in F]).astype(np.float32)
    arrays["omega_bands"] = omega.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "known_answer_temporal_filters.npz"), **arrays)
    print("temporal filters:", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    save_temporal_filters()
    t, r = synth.make_pair_u8(21, 1, 135, 240)   # image: ragged patches (38-pixel patches on 135x240), row-parity quirk levels
    save("feat_img_u8_135x240_fhd", t, r, "BCFHW", 0, "standard_fhd")
    t, r = synth.make_pair_u8(22, 6, 64, 100)    # video, 30 fps
    save("feat_vid_u8_6x64x100_fhd", t, r, "BCFHW", 30, "standard_fhd", "symmetric")
    t, r = synth.make_pair_u8(23, 1, 96, 160)    # 4K display: 76-pixel patches, batch of 2 via noise offset
    t2 = np.concatenate([t, np.clip(t.astype(np.int16) + 7, 0, 255).astype(np.uint8)], 0)
    save("feat_img_u8_b2_96x160_4k", t2, np.concatenate([r, r], 0), "BCFHW", 0, "standard_4k")
