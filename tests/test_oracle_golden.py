"""The numpy oracle against the reference-generated fixtures (no GPU needed)."""
import numpy as np
import pytest

import golden_util as gu
from oracle import cvvdp_oracle as O


@pytest.mark.parametrize("name", gu.case_names())
def test_oracle_matches_reference_fixture(name):
    z, meta = gu.load_case(name)
    jod, stats = O.predict(z["test"], z["ref"], meta["dim_order"], meta["fps"], meta["display"],
                           meta["padding"], meta["heatmap"], hm_block=meta.get("hm_block") or None)
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert np.allclose(stats["rho_band"], z["rho_band"])
    assert np.max(np.abs(np.asarray(jod, dtype=np.float64) - z["jod"])) <= gu.JOD_TOL
    if meta["heatmap"] == "raw":
        hm, hm_ref = stats["heatmap"].astype(np.float32), z["heatmap"].astype(np.float32)
        assert hm.shape == hm_ref.shape
        assert np.max(np.abs(hm - hm_ref)) <= gu.HEATMAP_ATOL
    elif meta["heatmap"] in ("threshold", "supra-threshold"):  # colour LUT x tone-mapped context, per block
        hm, hm_ref = stats["heatmap"].astype(np.float32), z["heatmap"].astype(np.float32)
        assert hm.shape == hm_ref.shape
        assert np.max(np.abs(hm - hm_ref)) <= gu.COLOR_HEATMAP_ATOL


@pytest.mark.parametrize("name", ["img_u8_70x121_4k", "vid_u8_10x64x100_fhd_rep"])
def test_oracle_stage_level(name):
    """Stage-level pins: DKL front end, temporal channels, Gaussian pyramid, contrast bands, log L_bkg."""
    z, meta = gu.load_case(name)
    dm, P = O.Display(meta["display"]), O.Params()
    T = O.reshuffle_dims(z["test"], meta["dim_order"])
    Rf = O.reshuffle_dims(z["ref"], meta["dim_order"])
    f = int(z["st_frame"])
    dkl_t, dkl_r = O.frontend(T[:, :, f], dm), O.frontend(Rf[:, :, f], dm)
    scale = np.abs(z["st_dkl_ref"]).max()
    assert np.max(np.abs(dkl_t - z["st_dkl_test"][:, :, 0])) <= 2e-4 * scale
    assert np.max(np.abs(dkl_r - z["st_dkl_ref"][:, :, 0])) <= 2e-4 * scale
    inter = {}
    O.predict(z["test"], z["ref"], meta["dim_order"], meta["fps"], meta["display"], meta["padding"],
              frame_range=(f, f + 1), intermediates=inter)
    R = inter["R"][f]
    R_ref = z["st_R"][:, :, 0]
    assert np.max(np.abs(R - R_ref)) <= 2e-4 * np.abs(R_ref).max()
    if meta["fps"] > 0:
        filt = O.temporal_filters(meta["fps"], P)
        assert np.max(np.abs(np.stack(filt) - z["st_filters"])) < 2e-6
    L = len(z["rho_band"])
    bands, logL, gpyr = O.weber_contrast_decompose(R_ref, L)
    for i in range(L):
        g_ref = z[f"st_gpyr{i}"][:, :, 0]
        assert gpyr[i].shape == g_ref.shape
        assert np.max(np.abs(gpyr[i] - g_ref)) <= 1e-5 * np.abs(g_ref).max()
        b_ref = z[f"st_band{i}"][:, :, 0]
        assert np.max(np.abs(bands[i] - b_ref)) <= 1e-4 * max(1.0, np.abs(b_ref).max())
        l_ref = z[f"st_logL{i}"][:, :, 0]
        assert np.max(np.abs(logL[i] - l_ref)) <= 1e-5


def test_identical_pair_is_10_jod():
    z, meta = gu.load_case("vid_u8_5x40x64_fhd_rep_short")
    jod, stats = O.predict(z["ref"], z["ref"], meta["dim_order"], meta["fps"], meta["display"])
    assert float(jod) == 10.0
    assert np.all(stats["Q_per_ch"] == 0)


def test_frame_range_matches_full_run():
    z, meta = gu.load_case("vid_u8_10x64x100_fhd_rep")
    _, full = O.predict(z["test"], z["ref"], meta["dim_order"], meta["fps"], meta["display"])
    _, part = O.predict(z["test"], z["ref"], meta["dim_order"], meta["fps"], meta["display"], frame_range=(4, 7))
    assert np.array_equal(part["Q_per_ch"][:, :, 4:7], full["Q_per_ch"][:, :, 4:7])
    assert np.all(part["Q_per_ch"][:, :, :4] == 0) and np.all(part["Q_per_ch"][:, :, 7:] == 0)


@pytest.mark.container
def test_known_answer_wavy_facade_blur():
    """examples/ex_simple_image.py:14-18: 'Blur - Quality: 8.514 JOD' (standard_4k)."""
    import cv2
    from scipy.ndimage import gaussian_filter
    im = cv2.imread("/root/reference/example_media/wavy_facade.png", cv2.IMREAD_UNCHANGED)[:, :, ::-1]
    blur = np.zeros_like(im)
    for cc in range(3):  # examples/ex_utils.py:27-41
        blur[..., cc] = gaussian_filter(im[..., cc], 2, mode="nearest", truncate=2.0)
    jod, _ = O.predict(blur, np.ascontiguousarray(im), "HWC", 0, "standard_4k")
    assert abs(float(jod) - 8.514) < 5e-4 + 1e-3


@pytest.mark.container
def test_oracle_vs_live_reference_random_shapes():
    """Cross-check against the live reference on a few extra odd shapes (container only)."""
    import torch
    from oracle import ref_loader
    pycvvdp = ref_loader.load()
    rng = np.random.default_rng(5)
    for (F, H, W, fps, disp, pad) in [(1, 33, 47, 0, "standard_fhd", "replicate"),
                                      (7, 37, 50, 50, "standard_phone", "symmetric"),
                                      (4, 64, 64, 120, "sdr_4k_30", "replicate")]:
        ref = (rng.random((1, 3, F, H, W)) * 255).astype(np.uint8)
        tst = np.clip(ref.astype(np.float32) + 6 * rng.standard_normal(ref.shape), 0, 255).astype(np.uint8)
        m = pycvvdp.cvvdp(display_name=disp, device=torch.device("cpu"), temp_padding=pad, quiet=True)
        with torch.no_grad():
            q, s = m.predict(tst, ref, dim_order="BCFHW", frames_per_second=fps)
        jod, stats = O.predict(tst, ref, "BCFHW", fps, disp, pad)
        gu.assert_q_close(stats["Q_per_ch"], s["Q_per_ch"], f"{F}x{H}x{W}")
        assert abs(float(jod) - float(q)) <= gu.JOD_TOL


@pytest.mark.parametrize("name", gu.yuv_case_names())
def test_oracle_yuv_ingestion(name, tmp_path):
    """Raw planar YUV files: unpack + chroma upsampling + matrix, then the whole path."""
    tf, rf, z, meta = gu.write_yuv_case(name, str(tmp_path))
    rgb, fps = O.read_yuv_rgb(tf)
    assert np.max(np.abs(rgb[0, :, -1].transpose(1, 2, 0) - z["rgb_last_test_frame"])) <= 2e-6
    jod, stats = O.predict_yuv(tf, rf, meta["display"], meta["padding"])
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL


@pytest.mark.parametrize("name", gu.vfile_case_names())
def test_oracle_video_file_sources(name, tmp_path):
    """Frames decoded by ffmpeg (planar YUV with the ffmpeg-path matrix, packed RGB of --ffmpeg-cc) and the full-screen
    resize, against fixtures generated with the reference's video_source_video_file /
    video_source_yuv_file."""
    tf, rf, z, meta = gu.write_vfile_case(name, str(tmp_path))
    oh, ow = z["rgb_first_test_frame"].shape[:2]
    rs = None if meta["full_screen_resize"] is None else (meta["full_screen_resize"], oh, ow)
    if meta["kind"] == "yuv":
        T, fps = O.read_yuv_rgb(tf, resize=rs)
        R, _ = O.read_yuv_rgb(rf, resize=rs)
    else:
        st = next(s for s in meta["probe"]["streams"] if s["codec_type"] == "video")
        T = O.read_ffmpeg_stream_rgb(tf, st, meta["ffmpeg_cc"], rs)
        R = O.read_ffmpeg_stream_rgb(rf, st, meta["ffmpeg_cc"], rs)
        fps = float(st["r_frame_rate"].split("/")[0])
    assert np.max(np.abs(T[0, :, 0].transpose(1, 2, 0) - z["rgb_first_test_frame"])) <= 5e-6
    jod, stats = O.predict(T, R, "BCFHW", fps, meta["display"], meta["padding"])
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL


@pytest.mark.parametrize("mode", ["nearest", "bilinear", "bicubic", "area"])
def test_oracle_resize_matches_torch(mode):
    import torch
    """The resize restatement against torch.nn.functional.interpolate itself (the third-party routine the reference
    calls), up- and down-scaling with non-integer factors."""
    g = torch.Generator().manual_seed(5)
    for (h, w, oh, ow) in [(36, 52, 72, 104), (36, 52, 50, 77), (48, 64, 30, 41), (37, 53, 37, 80), (20, 30, 61, 45)]:
        x = torch.rand((h, w, 3), generator=g) * 1.2 - 0.1
        want = torch.nn.functional.interpolate(x.permute(2, 0, 1)[None], size=(oh, ow), mode=mode)[0].permute(1, 2, 0).clip(0, 1)
        got = O.resize_rgb(x.numpy(), oh, ow, mode)
        assert np.max(np.abs(got - want.numpy())) <= 5e-6, (mode, h, w, oh, ow)


@pytest.mark.parametrize("name", gu.feature_case_names())
def test_oracle_features_match_reference(name):
    """per-band patch statistics of |T|S, |R|S, D (cvvdp_ml_metric.py:78-106, 302-352) against
    tensors produced by the reference's own cvvdp_ml_base.extract_features."""
    z, meta = gu.load_case(name)
    _, stats = O.predict(z["test"], z["ref"], meta["dim_order"], meta["fps"], meta["display"], meta["padding"], features=True)
    gu.assert_features_close(stats["features"], z, name)


def test_oracle_temporal_filters_match_reference():
    """Direct inverse real DFT of the oracle against the reference's irfft + fftshift (cvvdp_metric.py:1057-1092)."""
    import os
    z = np.load(os.path.join(gu.GOLDEN_DIR, "known_answer_temporal_filters.npz"))
    P = O.Params()
    for key in z.files:
        if key.startswith("fps_"):
            got = np.stack(O.temporal_filters(float(key[4:]), P))
            assert got.shape == z[key].shape and np.max(np.abs(got - z[key])) <= 2e-6, key


def test_oracle_prefiltered_source_fixture():
    """Pre-filtered video source (cvvdp_metric.py:470-488) against the reference-generated fixture."""
    z, meta = gu.load_case("prefilt_vid_f32_7x48x80_fhd")
    jod, stats = O.predict_prefiltered(z["test4"], z["ref4"], meta["fps"], meta["display"])
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], "prefiltered")
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL
