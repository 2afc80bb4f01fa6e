"""Synthetic planar YUV clips (limited range) derived from tests/synth.py -- shared by
tests/golden/make_golden_yuv.py and the GPU tests."""
import numpy as np

import synth


def synth_yuv(seed, F, H, W, chroma, bit_depth):
    """Planar YUV frames (limited range) built from the synthetic RGB clip: rough BT.709 forward matrix."""
    tst, ref = synth.make_pair_u8(seed, F, H, W)
    out = []
    for clip in (tst, ref):
        rgb = clip[0].astype(np.float32) / 255.0  # [3,F,H,W]
        Y = 0.2126 * rgb[0] + 0.7152 * rgb[1] + 0.0722 * rgb[2]
        Cb = (rgb[2] - Y) / 1.8556
        Cr = (rgb[0] - Y) / 1.5748
        sc = 2 ** (bit_depth - 8)
        planes = []
        for f in range(F):
            y = np.clip(np.rint((Y[f] * 219 + 16) * sc), 0, 255 * sc + sc - 1)
            cb, cr = (np.clip(np.rint((c[f] * 224 + 128) * sc), 0, 255 * sc + sc - 1) for c in (Cb, Cr))
            if chroma in ("420", "422"):
                cb, cr = (0.5 * (c[:, 0::2] + c[:, 1::2]) for c in (cb, cr))
            if chroma == "420":
                cb, cr = (0.5 * (c[0::2] + c[1::2]) for c in (cb, cr))
            planes += [y.ravel(), np.rint(cb).ravel(), np.rint(cr).ravel()]
        out.append(np.concatenate(planes).astype(np.uint16 if bit_depth > 8 else np.uint8))
    return out


