"""Full-shape parity against the UNMODIFIED reference running on the same B200 (`-m gpu`).

The reference (gfxdisp/ColorVideoVDP, `pycvvdp`) is installed by tools/stage_reference.sh into the git-ignored
baseline/_ref/, which travels to the GPU box with the snapshot; it runs on `cuda` with TF32 off -- "the reference
PyTorch path" of BASELINE.json's north_star -- and is the JOD / Q_per_ch oracle of the BASELINE configurations that
the numpy oracle is too slow for.  Tolerances: |dJOD| <= 1e-3 (north_star), |dQ_per_ch| <= 1e-3 |Q| + 1e-5,
raw heat map (fp16) <= 2e-3.  Skipped (loudly) when the reference has not been staged.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
import synth
from oracle import ref_loader as RL

import colorvideovdp_b200 as cv

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RL.staged(), reason="reference not staged: run tools/stage_reference.sh")]
DEV = torch.device("cuda:0")


def _clip(seed, F, H, W, hdr=False):
    """Synthetic pair generated on the device (bench.make_clip: same construction as tests/synth.py)."""
    import bench
    return bench.make_clip(seed, 0, F, H, W, "u16" if hdr else "u8", DEV, hdr=hdr)


def _reference(display, tst, ref, fps, heatmap=None):
    refm = RL.reference_metric(display, DEV, heatmap=heatmap)
    with torch.no_grad():
        jod, stats = refm.predict(tst, ref, dim_order="BCFHW", frames_per_second=fps)
    out = float(jod), np.asarray(stats["Q_per_ch"]), (stats["heatmap"] if heatmap else None)
    del refm
    torch.cuda.empty_cache()
    return out


def test_config2_1080p_60_frames_all_frames():
    """BASELINE configs[1]: 1920x1080, 60 frames, 30 fps, standard_fhd -- every frame, JOD and Q_per_ch."""
    tst, ref = _clip(2, 60, 1080, 1920)
    jod_r, Q_r, _ = _reference("standard_fhd", tst, ref, 30.0)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    jod, stats = m.predict(tst, ref, dim_order="BCFHW", frames_per_second=30.0)
    gu.assert_q_close(stats["Q_per_ch"], Q_r, "config 2")
    assert abs(float(jod) - jod_r) <= 1e-3, (float(jod), jod_r)
    # the host path (streamed upload) gives the same bits as the device path
    jod_h, stats_h = m.predict(tst.cpu(), ref.cpu(), dim_order="BCFHW", frames_per_second=30.0)
    assert np.array_equal(stats_h["Q_per_ch"], stats["Q_per_ch"])


def test_config3_4k_60fps_40_frames():
    """BASELINE configs[2] shape (3840x2160, 60 fps, standard_4k): 40 frames = 17-tap filter fully warmed up for
    23 of them; the reference needs ~40 GB for this clip on the GPU."""
    tst, ref = _clip(3, 40, 2160, 3840)
    jod_r, Q_r, _ = _reference("standard_4k", tst, ref, 60.0)
    m = cv.cvvdp(display_name="standard_4k", device=DEV)
    jod, stats = m.predict(tst, ref, dim_order="BCFHW", frames_per_second=60.0)
    gu.assert_q_close(stats["Q_per_ch"], Q_r, "config 3")
    assert abs(float(jod) - jod_r) <= 1e-3, (float(jod), jod_r)


def test_config4_hdr_pq_raw_heatmap_8_frames():
    """BASELINE configs[3] shape: 3840x2160 HDR (PQ code values as uint16) at 60 fps on standard_hdr_pq with the
    raw heat map, 8 frames: JOD, Q_per_ch and every heat-map pixel."""
    tst, ref = _clip(4, 8, 2160, 3840, hdr=True)
    jod_r, Q_r, hm_r = _reference("standard_hdr_pq", tst, ref, 60.0, heatmap="raw")
    m = cv.cvvdp(display_name="standard_hdr_pq", heatmap="raw", device=DEV)
    jod, stats = m.predict(tst, ref, dim_order="BCFHW", frames_per_second=60.0)
    gu.assert_q_close(stats["Q_per_ch"], Q_r, "config 4")
    assert abs(float(jod) - jod_r) <= 1e-3, (float(jod), jod_r)
    hm = stats["heatmap"]
    assert tuple(hm.shape) == tuple(hm_r.shape) and hm.dtype == torch.float16
    err = (hm.float() - hm_r.float().cpu()).abs().max().item()
    assert err <= gu.HEATMAP_ATOL, err


def test_config5_batch_items_match_reference_per_item():
    """BASELINE configs[4] per-item check on one GPU: a batch of two 4K clips (the per-rank share when 8 items run
    on 4 GPUs) through distributed.predict_sharded equals the reference's JOD of each item."""
    from colorvideovdp_b200 import distributed as D
    F, fps = 24, 60.0
    m = cv.cvvdp(display_name="standard_4k", device=DEV)
    pieces, jods_r = [], []
    for item in range(2):
        tst, ref = _clip(3 + 17 * item, F, 2160, 3840)
        jod_r, _, _ = _reference("standard_4k", tst, ref, fps)
        jods_r.append(jod_r)
        pieces.append((item, 0, F, 0, tst, ref))
    jod, Q = D.predict_sharded(m, pieces, 2, F, fps)
    assert np.max(np.abs(jod.cpu().numpy() - np.asarray(jods_r))) <= 1e-3, (jod, jods_r)


def test_reference_cuda_vs_cpu_noise_floor():
    """The reference disagrees with itself between cuda and cpu by this much (context for the 1e-3 gates)."""
    tst, ref = synth.make_pair_u8(9, 4, 270, 480)
    t, r = torch.from_numpy(tst), torch.from_numpy(ref)
    jg, Qg, _ = _reference("standard_4k", t.to(DEV), r.to(DEV), 30.0)
    refc = RL.reference_metric("standard_4k", "cpu")
    with torch.no_grad():
        jc, sc = refc.predict(t, r, dim_order="BCFHW", frames_per_second=30.0)
    assert abs(jg - float(jc)) <= 1e-3
    gu.assert_q_close(Qg, np.asarray(sc["Q_per_ch"]), "reference cuda vs cpu")
