"""Fixture for pre-filtered video sources (cvvdp_metric.py:470-488), generated from the UNMODIFIED reference.
Run in the build container only:  python tests/golden/make_golden_prefiltered.py

A tiny third-party `video_source` with `is_temporally_filtered = True` hands the reference four-channel
'DKLd65_trans' frames (here: DKL of a synthetic clip plus a crude transient channel); the fixture stores those
frames and the reference's Q_per_ch / JOD, so the test feeds the same frames to the B200 path and the oracle."""
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
from oracle import cvvdp_oracle as O  # noqa: E402
import synth  # noqa: E402

pycvvdp = ref_loader.load(prefer_staged=False)


def four_channels(clip_u8, dm):
    """[1,3,F,H,W] uint8 -> [1,4,F,H,W] fp32: DKL of every frame + frame difference of the achromatic channel."""
    F = clip_u8.shape[2]
    dkl = np.stack([O.frontend(clip_u8[:, :, f], dm) for f in range(F)], axis=2)  # [1,3,F,H,W]
    trans = np.zeros_like(dkl[:, :1])
    trans[:, 0, 1:] = dkl[:, 0, 1:] - dkl[:, 0, :-1]
    return np.concatenate([dkl, trans], axis=1).astype(np.float32)


class Prefiltered(pycvvdp.video_source.video_source):
    is_temporally_filtered = True

    def __init__(self, t4, r4, fps):
        self.t4, self.r4, self.fps = torch.from_numpy(t4), torch.from_numpy(r4), fps

    def get_video_size(self):
        return (self.t4.shape[3], self.t4.shape[4], self.t4.shape[2])

    def get_frames_per_second(self):
        return self.fps

    def get_test_frame(self, frame, device, colorspace):
        assert colorspace == "DKLd65_trans"
        return self.t4[:, :, frame:frame + 1].to(device)

    def get_reference_frame(self, frame, device, colorspace):
        assert colorspace == "DKLd65_trans"
        return self.r4[:, :, frame:frame + 1].to(device)


if __name__ == "__main__":
    tst, ref = synth.make_pair_u8(93, 7, 48, 80)
    dm = O.Display("standard_fhd")
    t4, r4 = four_channels(tst, dm), four_channels(ref, dm)
    m = pycvvdp.cvvdp(display_name="standard_fhd", device=torch.device("cpu"), quiet=True)
    with torch.no_grad():
        q, s = m.predict_video_source(Prefiltered(t4, r4, 30))
    meta = {"fps": 30, "display": "standard_fhd",
            "reference": "gfxdisp/ColorVideoVDP pycvvdp 0.5.4 (params 0.5.6), torch %s CPU" % torch.__version__}
    np.savez_compressed(os.path.join(HERE, "prefilt_vid_f32_7x48x80_fhd.npz"), test4=t4, ref4=r4,
                        meta=np.asarray(json.dumps(meta)), jod=np.asarray(q.cpu().numpy(), dtype=np.float32),
                        Q_per_ch=s["Q_per_ch"].astype(np.float32))
    print("prefilt fixture: JOD", float(q), s["Q_per_ch"].shape)
