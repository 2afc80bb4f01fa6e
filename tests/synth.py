"""Seeded synthetic test/reference clips.

Generated on the CPU with numpy's PCG64 so the byte stream does not depend on torch:
reference = low-pass random field + translating sinusoidal grating + 2 % per-pixel noise, scaled to
16..235; test = reference + per-frame Gaussian noise (sigma = 4 levels) + a +3-level offset on red.
Used by tests/ (small sizes) and bench.py (BASELINE.json configs).
"""
import numpy as np


def _lowpass_field(rng, F, H, W):
    h8, w8 = max(H // 8, 2), max(W // 8, 2)
    f = rng.random((h8 + 2, w8 + 2), dtype=np.float32)
    for _ in range(3):  # 3x box filter
        f = (f[:-1, :-1] + f[1:, :-1] + f[:-1, 1:] + f[1:, 1:]) * 0.25
        f = np.pad(f, ((0, 1), (0, 1)), mode="edge")
    # bilinear upsample to HxW
    ys = np.linspace(0, f.shape[0] - 1.001, H, dtype=np.float32)
    xs = np.linspace(0, f.shape[1] - 1.001, W, dtype=np.float32)
    y0, x0 = ys.astype(np.int32), xs.astype(np.int32)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    a = f[y0][:, x0] * (1 - fx) + f[y0][:, x0 + 1] * fx
    b = f[y0 + 1][:, x0] * (1 - fx) + f[y0 + 1][:, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def make_pair_u8(seed, F, H, W, C=3, noise_sigma=4.0, red_offset=3.0):
    """Returns (test, ref) uint8 arrays of shape [1, C, F, H, W]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ref = np.empty((1, C, F, H, W), dtype=np.uint8)
    tst = np.empty((1, C, F, H, W), dtype=np.uint8)
    xx = np.arange(W, dtype=np.float32)[None, :]
    yy = np.arange(H, dtype=np.float32)[:, None]
    base = [_lowpass_field(rng, F, H, W) for _ in range(C)]
    for f in range(F):
        for c in range(C):
            grating = 0.5 + 0.5 * np.sin((xx + 0.5 * yy - 2.0 * f) * (2 * np.pi / (16.0 + 8 * c)))
            img = 0.55 * base[c] + 0.35 * grating + 0.02 * rng.standard_normal((H, W), dtype=np.float32) + 0.05
            r = np.clip(16.0 + img * 219.0, 16, 235)
            t = r + noise_sigma * rng.standard_normal((H, W), dtype=np.float32) + (red_offset if c == 0 else 0.0)
            ref[0, c, f] = np.clip(np.rint(r), 0, 255).astype(np.uint8)
            tst[0, c, f] = np.clip(np.rint(t), 0, 255).astype(np.uint8)
    return tst, ref


def make_pair_pq_u16(seed, F, H, W):
    """HDR: 10-bit PQ code values 64..800 stored as uint16 (code * 64)."""
    t8, r8 = make_pair_u8(seed, F, H, W)
    def conv(a):
        code = 64.0 + (a.astype(np.float32) - 16.0) / 219.0 * (800.0 - 64.0)
        return (np.clip(np.rint(code), 0, 1023).astype(np.uint16) * 64).astype(np.uint16)
    return conv(t8), conv(r8)
