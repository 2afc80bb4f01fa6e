"""Kernel + host logic on the mock device (tests/emu): the *same* CUDA sources compiled for the CPU,
checked against the reference-generated fixtures and the numpy oracle.  This is what keeps tile/halo
indexing, padding rules and the host orchestration honest on a machine without a GPU; numerics of
the real MUFU paths are covered by the -m gpu tests."""
import os

import numpy as np
import pytest
import torch

import golden_util as gu
import synth
from emu_util import mock_device  # noqa: F401
from oracle import cvvdp_oracle as O

import colorvideovdp_b200 as cv


@pytest.mark.parametrize("name", gu.case_names())
def test_golden_cases_on_mock_device(name, mock_device):
    z, meta = gu.load_case(name)
    # hm_block == 1: the reference coloured this clip frame by frame (its CPU block size); one frame per pass here too
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], heatmap=meta["heatmap"],
                 gpu_mem=1e-6 if meta.get("hm_block") == 1 else None)
    jod, stats = m.predict(z["test"], z["ref"], dim_order=meta["dim_order"], frames_per_second=meta["fps"])
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert np.max(np.abs(np.asarray(jod, dtype=np.float64) - z["jod"])) <= gu.JOD_TOL
    assert np.allclose(stats["rho_band"], z["rho_band"])
    assert stats["N_frames"] == z["Q_per_ch"].shape[2]
    if meta["heatmap"] == "raw":
        hm = stats["heatmap"]
        assert hm.dtype == torch.float16 and tuple(hm.shape) == z["heatmap"].shape
        assert np.max(np.abs(hm.float().numpy() - z["heatmap"].astype(np.float32))) <= gu.HEATMAP_ATOL
    elif meta["heatmap"] in ("threshold", "supra-threshold"):  # colour LUT x tone-mapped context image
        hm = stats["heatmap"]
        assert hm.dtype == torch.float16 and tuple(hm.shape) == z["heatmap"].shape
        assert np.max(np.abs(hm.float().numpy() - z["heatmap"].astype(np.float32))) <= gu.COLOR_HEATMAP_ATOL


def test_identical_pair_is_exactly_10(mock_device):
    tst, ref = synth.make_pair_u8(3, 4, 40, 56)
    m = cv.cvvdp(display_name="standard_fhd")
    jod, stats = m.predict(ref, ref, frames_per_second=30)
    assert float(jod) == 10.0
    assert np.all(stats["Q_per_ch"] == 0)


@pytest.mark.parametrize("shape,fps,heatmap", [((1, 150, 360), 0, None), ((1, 131, 250), 0, "raw"), ((5, 64, 236), 30, None)])
def test_band_kernel_strips_segments_and_packed_temporal_kernel(shape, fps, heatmap, mock_device):
    """The band kernel on levels several 48-column strips wide (several row segments, a ragged last strip, phase C
    trailing by 16 rows across the segment ends, the heat-map variant); clips whose planes are whole 64-pixel warp
    segments take the packed two-stage temporal kernel.  Checked against the oracle."""
    F, H, W = shape
    tst, ref = synth.make_pair_u8(11, F, H, W)
    m = cv.cvvdp(display_name="standard_fhd", heatmap=heatmap)
    jod, stats = m.predict(tst, ref, frames_per_second=fps)
    assert m._ctx.band_strip_width(0) == 48
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, "standard_fhd", heatmap=heatmap)
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], str(shape))
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL
    if heatmap:
        assert np.max(np.abs(stats["heatmap"].float().numpy() - stats_o["heatmap"].astype(np.float32))) <= gu.HEATMAP_ATOL


@pytest.mark.parametrize("shape", [(6, 16, 64), (3, 20, 28)])  # two-stage temporal kernel / generic kernel
def test_input_validation_on_the_fused_path(shape, mock_device, caplog):
    """The reference's per-frame input checks (display_model.py:335-337, video_source.py:48-72,
    cvvdp_metric.py:906-907) from the device counters of the fused front end: out-of-range values warn and are
    clamped, a NaN warns and fails with the reference's assertion, nearly black content warns about the scale."""
    import logging
    F, H, W = shape
    tst, ref = synth.make_pair_u8(21, F, H, W)
    tf, rf = tst.astype(np.float32) / 255, ref.astype(np.float32) / 255
    m = cv.cvvdp(display_name="standard_fhd")
    with caplog.at_level(logging.WARNING):
        j_clean, _ = m.predict(tf, rf, frames_per_second=30)
    assert not caplog.records
    hot = tf.copy()
    hot[0, 1, 2, 5, 7] = 1.2
    clamped = hot.copy()
    clamped[0, 1, 2, 5, 7] = 1.0
    with caplog.at_level(logging.WARNING):
        j_hot, s_hot = m.predict(hot, rf, frames_per_second=30)
    assert any("Pixel outside the valid range 0-1" in r.message for r in caplog.records)
    j_cl, s_cl = m.predict(clamped, rf, frames_per_second=30)
    assert np.array_equal(s_hot["Q_per_ch"], s_cl["Q_per_ch"])  # clamped, exactly like the reference
    caplog.clear()
    bad = tf.copy()
    bad[0, 0, 1, 3, 3] = np.nan
    with caplog.at_level(logging.WARNING):
        with pytest.raises(AssertionError, match="Must not be nan"):
            m.predict(bad, rf, frames_per_second=30)
    assert any("NaN" in r.message for r in caplog.records)
    caplog.clear()
    with caplog.at_level(logging.WARNING):  # the counters were reset: a clean clip is clean again
        j_again, _ = m.predict(tf, rf, frames_per_second=30)
    assert not caplog.records and float(j_again) == float(j_clean)
    dark = cv.cvvdp(display_name="standard_fhd", display_photometry=cv.vvdp_display_photo_eotf(0.5, contrast=1000, EOTF="sRGB", E_ambient=0))
    with caplog.at_level(logging.WARNING):
        dark.predict(tf, rf, frames_per_second=30)
    assert any("mean color value is less than 1" in r.message for r in caplog.records)


def test_frame_blocks_and_ranges_are_partition_independent(mock_device):
    """Blocking (gpu_mem limit -> 1..n frames per pass) and frame ranges must not change any value."""
    tst, ref = synth.make_pair_u8(4, 9, 36, 48)
    m = cv.cvvdp(display_name="standard_fhd", temp_padding="symmetric")
    _, full = m.predict(tst, ref, frames_per_second=24)
    small = cv.cvvdp(display_name="standard_fhd", temp_padding="symmetric", gpu_mem=1e-6)  # -> 1 frame per block
    _, blk = small.predict(tst, ref, frames_per_second=24)
    assert small._info.block_frames == 1
    assert np.array_equal(full["Q_per_ch"], blk["Q_per_ch"])
    vs = cv.video_source_array(tst, ref, 24, display_photometry=m.display_photometry)
    Qa, _ = m.compute_q_per_ch(vs, (0, 4))
    Qb, _ = m.compute_q_per_ch(vs, (4, 9))
    assert np.array_equal((Qa + Qb).numpy(), full["Q_per_ch"])


def test_dim_orders_and_strided_views(mock_device):
    tst, ref = synth.make_pair_u8(5, 1, 33, 47)
    m = cv.cvvdp(display_name="standard_4k")
    j0, s0 = m.predict(tst, ref, dim_order="BCFHW")
    hwc_t = np.ascontiguousarray(tst[0, :, 0].transpose(1, 2, 0))
    hwc_r = np.ascontiguousarray(ref[0, :, 0].transpose(1, 2, 0))
    j1, s1 = m.predict(hwc_t, hwc_r, dim_order="HWC")
    whc_t, whc_r = np.ascontiguousarray(hwc_t.transpose(1, 0, 2)), np.ascontiguousarray(hwc_r.transpose(1, 0, 2))
    j2, s2 = m.predict(whc_t, whc_r, dim_order="WHC")
    assert np.array_equal(s0["Q_per_ch"], s1["Q_per_ch"]) and np.array_equal(s0["Q_per_ch"], s2["Q_per_ch"])
    assert float(j0) == float(j1) == float(j2)


@pytest.mark.parametrize("H,W", [(36, 52), (32, 64)])  # (32, 64): whole warp segments -> two-stage temporal kernel on DKL input
def test_plugin_video_source_matches_fast_path(H, W, mock_device):
    """A third-party video_source (frames pulled one by one in DKLd65) must agree with the fused path."""
    tst, ref = synth.make_pair_u8(6, 7, H, W)
    m = cv.cvvdp(display_name="standard_fhd")
    _, fast = m.predict(tst, ref, frames_per_second=30)
    dm, P = O.Display("standard_fhd"), None

    class OracleFrontendSource(cv.video_source):
        def get_video_size(self):
            return (H, W, 7)

        def get_frames_per_second(self):
            return 30

        def _frame(self, arr, f, device):
            return torch.from_numpy(O.frontend(arr[:, :, f], dm))[:, :, None].to(device)

        def get_test_frame(self, f, device, colorspace):
            assert colorspace == "DKLd65"
            return self._frame(tst, f, device)

        def get_reference_frame(self, f, device, colorspace):
            return self._frame(ref, f, device)

    _, plug = m.predict_video_source(OracleFrontendSource())
    gu.assert_q_close(plug["Q_per_ch"], fast["Q_per_ch"], "plugin vs fast")


def test_display_model_forward_on_mock_device(mock_device):
    """display_model plugin surface: forward() and source_2_target_colorspace('DKLd65')."""
    rng = np.random.default_rng(0)
    V = rng.random((2, 3, 1, 20, 24), dtype=np.float32)
    for name in ("standard_4k", "standard_hdr_pq", "standard_hdr_linear", "standard_hdr_hlg"):
        dm = cv.vvdp_display_photometry.load(name, [])
        odm = O.Display(name)
        # the mock device has no CUDA: call the kernel wrapper directly on CPU tensors
        from colorvideovdp_b200 import _native as N
        from colorvideovdp_b200.cvvdp_metric import _default_native_inputs
        import emu_util
        ctx = N.Context(*_default_native_inputs(), 0, library=emu_util.emu_library())
        ctx.set_display(dm.native_display())
        Vt = torch.from_numpy(V)
        out = torch.empty((2, 3, 20, 24))
        clip = N.Clip()
        clip.data = Vt.data_ptr()
        for i, s in enumerate(Vt.stride()):
            clip.stride[i] = s
        clip.frame0, clip.n_frames = 0, 1
        ctx.frontend(clip, 2, 3, 20, 24, N.DTYPE_F32, 0, N.CS_RGB_LINEAR, out.data_ptr(), None, None)
        L_ref = O.eotf_forward(V[:, :, 0], odm)
        assert np.max(np.abs(out.numpy() - L_ref) / np.abs(L_ref)) < 1e-4, name
        ctx.frontend(clip, 2, 3, 20, 24, N.DTYPE_F32, 0, N.CS_DKLD65, out.data_ptr(), None, None)
        D_ref = O.frontend(V[:, :, 0], odm)
        assert np.max(np.abs(out.numpy() - D_ref)) < 1e-4 * np.abs(D_ref).max(), name
        ctx.close()


def test_pooling_entry_point_matches_oracle(mock_device):
    rng = np.random.default_rng(1)
    m = cv.cvvdp(display_name="standard_4k")
    P = O.Params()
    for shape in [(1, 3, 1, 8), (2, 4, 17, 9), (3, 4, 300, 5)]:
        Q = (rng.random(shape) * 2).astype(np.float32)
        Q[0, 0, 0, 0] = 0
        jod = m.do_pooling_and_jods(Q)
        ref = O.do_pooling_and_jods(Q, P)
        assert np.allclose(np.atleast_1d(jod.numpy()), ref, atol=2e-5)


def test_error_behaviour(mock_device):
    m = cv.cvvdp(display_name="standard_4k")
    a = np.zeros((1, 3, 4, 32, 32), np.uint8)
    with pytest.raises(RuntimeError, match="frames_per_second"):
        m.predict(a, a)  # video without fps (video_source.py:279-280)
    with pytest.raises(RuntimeError, match="1 or 3 color channels"):
        m.predict(np.zeros((1, 2, 1, 32, 32), np.uint8), np.zeros((1, 2, 1, 32, 32), np.uint8))
    with pytest.raises(RuntimeError, match="same shape"):
        m.predict(np.zeros((1, 3, 1, 32, 32), np.uint8), np.zeros((1, 3, 1, 32, 40), np.uint8))
    hm = cv.cvvdp(display_name="standard_4k", heatmap="raw")
    with pytest.raises(cv.vq_exception):
        hm.predict(np.zeros((2, 3, 1, 32, 32), np.uint8), np.zeros((2, 3, 1, 32, 32), np.uint8))
    bad = cv.cvvdp(display_name="standard_4k", temp_padding="valid")
    with pytest.raises(RuntimeError, match="padding"):
        bad.predict(a, a, frames_per_second=30)
    with pytest.raises(RuntimeError, match="Display model not found"):
        cv.cvvdp(display_name="no_such_display")


def test_host_streaming_chunks_match_resident_path(mock_device):
    """process_host uploads in chunks and moves the history frames device-to-device between its two
    staging buffers; the result must be bit-identical to the resident path (process_device)."""
    tst, ref = synth.make_pair_u8(8, 40, 24, 40)
    for padding in ("replicate", "symmetric"):
        m = cv.cvvdp(display_name="standard_fhd", temp_padding=padding)
        t, r = torch.from_numpy(tst), torch.from_numpy(ref)
        Qh, _ = m.q_per_ch_from_tensors(t, r, 40, 30, _resident=False)   # 3 chunks of 16 frames
        Qd, _ = m.q_per_ch_from_tensors(t, r, 40, 30, _resident=True)
        assert torch.equal(Qh, Qd)
        # a frame shard from a window of the clip, host path
        Qw, _ = m.q_per_ch_from_tensors(t[:, :, 10:40], r[:, :, 10:40], 40, 30, (20, 40), 10, _resident=False)
        assert torch.equal(Qw[:, :, 20:], Qd[:, :, 20:]) and bool((Qw[:, :, :20] == 0).all())
    # non-BCFHW host layout (frames outermost): FCHW view
    fchw_t = np.ascontiguousarray(tst[0].transpose(1, 0, 2, 3))
    fchw_r = np.ascontiguousarray(ref[0].transpose(1, 0, 2, 3))
    m = cv.cvvdp(display_name="standard_fhd")
    _, s1 = m.predict(fchw_t, fchw_r, dim_order="FCHW", frames_per_second=30)
    _, s2 = m.predict(tst, ref, frames_per_second=30)
    assert np.array_equal(s1["Q_per_ch"], s2["Q_per_ch"])


@pytest.mark.parametrize("name", gu.yuv_case_names())
def test_yuv_files_on_mock_device(name, tmp_path, mock_device):
    """video_source_yuv_file: fused YUV front end (fast path), the single-frame plugin surface and the
    generic plugin path must all match the reference-generated fixture."""
    tf, rf, z, meta = gu.write_yuv_case(name, str(tmp_path))
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"])
    vs = cv.video_source_yuv_file(tf, rf, display_photometry=meta["display"])
    jod, stats = m.predict_video_source(vs)
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL
    F = z["Q_per_ch"].shape[2]
    rgb = vs.test_vidr.get_frame_rgb_tensor(F - 1, torch.device("cpu"))
    assert tuple(rgb.shape) == z["rgb_last_test_frame"].shape
    assert np.max(np.abs(rgb.numpy() - z["rgb_last_test_frame"])) <= 2e-6

    class Wrapped(cv.video_source):  # forces the generic plugin path (frames pulled one by one in DKLd65)
        def get_video_size(self):
            return vs.get_video_size()

        def get_frames_per_second(self):
            return vs.get_frames_per_second()

        def get_test_frame(self, f, device, colorspace):
            return vs.get_test_frame(f, device, colorspace)

        def get_reference_frame(self, f, device, colorspace):
            return vs.get_reference_frame(f, device, colorspace)

    _, plug = m.predict_video_source(Wrapped())
    gu.assert_q_close(plug["Q_per_ch"], z["Q_per_ch"], name + " (plugin path)")


def open_vfile_source(tf, rf, meta, **extra):
    """The video source a vfile_* fixture was generated with."""
    if meta["kind"] == "yuv":
        return cv.video_source_yuv_file(tf, rf, display_photometry=meta["display"],
                                        full_screen_resize=meta["full_screen_resize"],
                                        resize_resolution=tuple(meta["resize_resolution"]),
                                        retain_aspect_ratio=meta["retain_aspect_ratio"])
    rr = meta["resize_resolution"]
    return cv.video_source_video_file(tf, rf, display_photometry=meta["display"], full_screen_resize=meta["full_screen_resize"],
                                      resize_resolution=None if rr is None else tuple(rr), ffmpeg_cc=meta["ffmpeg_cc"], **extra)


@pytest.mark.parametrize("name", gu.vfile_case_names())
def test_video_file_sources_on_mock_device(name, tmp_path, mock_device, monkeypatch):
    """ffmpeg-pipe readers (planar YUV decoded on the device, --ffmpeg-cc packed RGB) and the full-screen resize against
    fixtures generated with the reference's video_source_video_file / video_source_yuv_file."""
    monkeypatch.setenv("PATH", gu.FAKE_FFMPEG_DIR + os.pathsep + os.environ["PATH"])
    tf, rf, z, meta = gu.write_vfile_case(name, str(tmp_path))
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"])
    vs = open_vfile_source(tf, rf, meta)
    fused = []
    run_yuv = m._run_yuv
    monkeypatch.setattr(m, "_run_yuv", lambda *a, **k: (fused.append(1), run_yuv(*a, **k))[1])
    jod, stats = m.predict_video_source(vs)
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL
    # planar YUV without a resize goes through the fused temporal kernel; everything else frame by frame
    expect_fused = meta["full_screen_resize"] is None and not meta.get("ffmpeg_cc", False)
    assert bool(fused) == expect_fused
    H, W, F = vs.get_video_size()
    assert (stats["height"], stats["width"], stats["N_frames"]) == (H, W, F) == z["rgb_first_test_frame"].shape[:2] + (z["Q_per_ch"].shape[2],)
    # the decoded (and resized) first frame, through the reader interface of the reference
    if meta["kind"] == "yuv":
        rd = cv.video_reader_yuv(tf, resize_fn=meta["full_screen_resize"], resize_height=H, resize_width=W)
    else:
        vs2 = open_vfile_source(tf, rf, meta)
        vs2.init_readers()
        rd = vs2.test_vidr
    rgb = rd.unpack(rd.get_frame(), torch.device("cpu"))
    assert tuple(rgb.shape) == z["rgb_first_test_frame"].shape
    assert np.max(np.abs(rgb.numpy() - z["rgb_first_test_frame"])) <= 5e-6
    rd.close()
    if expect_fused:  # the frame-by-frame path of the same source gives the same answer
        vs3 = open_vfile_source(tf, rf, meta)
        monkeypatch.setattr(vs3, "yuv_readers", lambda: None)
        _, plug = m.predict_video_source(vs3)
        gu.assert_q_close(plug["Q_per_ch"], z["Q_per_ch"], name + " (frame by frame)")


def test_yuv_path_walks_the_clip_in_bounded_windows(tmp_path, mock_device, monkeypatch):
    """Host windows of the fused YUV path: a small window budget (one plan block per process call, pipes read
    forward only) gives exactly the result of a single window."""
    monkeypatch.setenv("PATH", gu.FAKE_FFMPEG_DIR + os.pathsep + os.environ["PATH"])
    name = "vfile_pipe_444_12b_unknown_12x40x56_sym"
    tf, rf, z, meta = gu.write_vfile_case(name, str(tmp_path))
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], gpu_mem=1e-6)  # one frame per pass
    _, whole = m.predict_video_source(open_vfile_source(tf, rf, meta))
    calls = []
    ph = m._ctx.process_host
    monkeypatch.setattr(m._ctx, "process_host", lambda t, r, f0, f1, *a: (calls.append((f0, f1, t.frame0, t.n_frames)), ph(t, r, f0, f1, *a))[1])
    monkeypatch.setattr(m, "yuv_chunk_bytes", 1)
    _, parts = m.predict_video_source(open_vfile_source(tf, rf, meta))
    assert len(calls) > 1 and calls[0][0] == 0 and calls[-1][1] == 12
    assert all(c[2] <= c[0] for c in calls)
    assert np.array_equal(parts["Q_per_ch"], whole["Q_per_ch"])
    gu.assert_q_close(parts["Q_per_ch"], z["Q_per_ch"], name)


def test_temporal_resampling_of_file_sources(tmp_path, mock_device, monkeypatch):
    """video_source_temp_resample_file (--temp-resample): a 30 fps test against a 60 fps reference, frames repeated to
    the common rate, against the reference's own class."""
    monkeypatch.setenv("PATH", gu.FAKE_FFMPEG_DIR + os.pathsep + os.environ["PATH"])
    tf, rf, z, meta = gu.write_vfile_case("vresample_30_vs_60fps_8x40x56", str(tmp_path))
    with pytest.raises(cv.vq_exception, match="different frame rates"):
        cv.video_source_video_file(tf, rf, display_photometry=meta["display"]).init_readers()
    vs = cv.video_source_temp_resample_file(tf, rf, display_photometry=meta["display"])
    assert vs.get_frames_per_second() == meta["resample_fps"] and list(vs.get_video_size()) == meta["video_size"]
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"])
    jod, stats = m.predict_video_source(vs)
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], "temporal resampling")
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL


def test_video_file_source_errors(tmp_path, mock_device, monkeypatch):
    """Error behaviour of the file sources (video_source_file.py:76-88, 455-465)."""
    with pytest.raises(cv.vq_exception, match="not found"):
        cv.video_source_video_file(str(tmp_path / "a.mp4"), str(tmp_path / "b.mp4")).init_readers()
    monkeypatch.setenv("PATH", gu.FAKE_FFMPEG_DIR + os.pathsep + os.environ["PATH"])
    (tmp_path / "broken.mp4").write_bytes(b"xx")
    with pytest.raises(cv.vq_exception, match="ffmpeg failed to open file"):
        cv.video_reader(str(tmp_path / "broken.mp4"))
    tf, rf, z, meta = gu.write_vfile_case("vfile_pipe_420_8b_bt709_5x48x64", str(tmp_path))
    vs = open_vfile_source(tf, rf, meta)
    vs.get_reference_frame(0, torch.device("cpu"), "DKLd65")
    with pytest.raises(cv.vq_exception, match="frame-by-frame"):
        vs.get_reference_frame(2, torch.device("cpu"), "DKLd65")
    assert vs.yuv_readers() is None  # a source that was already read is not handed to the fused path
    vs.reference_vidr.close(), vs.test_vidr.close()
    short = open_vfile_source(tf, rf, meta, frames=3)
    assert short.get_video_size() == (48, 64, 3)


@pytest.mark.parametrize("mode", ["nearest", "bilinear", "bicubic", "area"])
def test_resize_kernel_matches_torch_interpolate(mode, mock_device):
    """k_resize against torch.nn.functional.interpolate (what the reference calls for --full-screen-resize)."""
    from colorvideovdp_b200 import _native as N
    from colorvideovdp_b200 import cvvdp_metric as cm
    params, lut = cm._default_native_inputs()
    ctx = N.Context(params, lut, 0, library=cm._mock_library)
    g = torch.Generator().manual_seed(3)
    for (H, W, OH, OW) in [(36, 52, 72, 104), (36, 52, 50, 77), (48, 64, 30, 41), (37, 53, 37, 80), (64, 96, 16, 24), (20, 30, 61, 45)]:
        src = torch.rand((3, H, W), generator=g) * 1.2 - 0.1
        dst = torch.empty((3, OH, OW))
        ctx.resize(src.data_ptr(), dst.data_ptr(), 3, H, W, OH, OW, mode, True, None)
        want = torch.nn.functional.interpolate(src[None], size=(OH, OW), mode=mode)[0].clip(0, 1)
        assert float((dst - want).abs().max()) <= 5e-6, (mode, H, W, OH, OW)
    with pytest.raises(N.NativeError):
        ctx.resize(src.data_ptr(), dst.data_ptr(), 3, H, W, OH, OW, "lanczos", True, None)


@pytest.mark.parametrize("chroma,bit_depth,color_space,display", [("420", 8, "709", "standard_fhd"), ("422", 10, "2020", "standard_hdr_pq"),
                                                               ("444", 8, "709", "standard_hdr_hlg"), ("420", 10, "2020", "standard_hdr_pq")])
def test_yuv_two_stage_front_end(chroma, bit_depth, color_space, display, tmp_path, mock_device, monkeypatch):
    """Rows of whole 64-pixel segments take the packed two-stage temporal kernel with the planar-YUV front end (direct
    chroma taps); it must agree with the oracle and with the frame-by-frame path (k_frontend + generic kernel)."""
    from golden.make_golden_yuv_synth import synth_yuv
    F, H, W = 6, 34, 128
    t, r = synth_yuv(91, F, H, W, chroma, bit_depth)
    props = {"width": W, "height": H, "fps": 30, "bit_depth": bit_depth, "color_space": color_space, "chroma_ss": chroma}
    tf, rf = str(tmp_path / cv.create_yuv_fname("t", props)), str(tmp_path / cv.create_yuv_fname("r", props))
    t.tofile(tf), r.tofile(rf)
    m = cv.cvvdp(display_name=display)
    jod, fast = m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry=display))
    jod_o, want = O.predict_yuv(tf, rf, display)
    gu.assert_q_close(fast["Q_per_ch"], want["Q_per_ch"], "two-stage YUV vs oracle")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL
    vs = cv.video_source_yuv_file(tf, rf, display_photometry=display)
    monkeypatch.setattr(vs, "yuv_readers", lambda: None)
    _, slow = m.predict_video_source(vs)
    gu.assert_q_close(fast["Q_per_ch"], slow["Q_per_ch"], "two-stage YUV vs frame by frame")


def test_yuv_files_are_read_through_descriptors(tmp_path, mock_device, monkeypatch):
    """.yuv pairs reach the library as file descriptors (cvvdp_b200_process_files: pread into the staging slots, no
    mapping); several windows, an offset into the clip, and a file that ends early."""
    from colorvideovdp_b200 import _native as N
    name = "yuv_422_8b_709_12x40x48_sym"
    tf, rf, z, meta = gu.write_yuv_case(name, str(tmp_path))
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], gpu_mem=1e-6)
    calls = []
    pf = m._ctx.process_files
    monkeypatch.setattr(m._ctx, "process_files", lambda *a: (calls.append(a[2:8]), pf(*a))[1])
    monkeypatch.setattr(m, "yuv_chunk_bytes", 1)
    _, stats = m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry=meta["display"]))
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert len(calls) > 1 and calls[0][4:] == (0, calls[0][5]) and calls[-1][5] == 12
    # frames 4.. of the clip as a clip of their own == the same frames cut out of the files
    vs = cv.video_source_yuv_file(tf, rf, display_photometry=meta["display"])
    vs.set_offset(4)
    vs.set_num_frames(8)
    _, tail = m.predict_video_source(vs)
    fb = vs.test_vidr.frame_bytes
    tf2, rf2 = str(tmp_path / "cut" / os.path.basename(tf)), str(tmp_path / "cut" / os.path.basename(rf))
    os.makedirs(str(tmp_path / "cut"))
    for src, dst in ((tf, tf2), (rf, rf2)):
        with open(src, "rb") as f, open(dst, "wb") as g:
            g.write(f.read()[4 * fb:])
    _, cut = m.predict_video_source(cv.video_source_yuv_file(tf2, rf2, display_photometry=meta["display"]))
    assert np.array_equal(tail["Q_per_ch"], cut["Q_per_ch"])
    # a reference file that loses its last bytes after the reader counted its frames
    vs = cv.video_source_yuv_file(tf, rf, display_photometry=meta["display"])
    vs.reference_vidr.fileno()
    with open(rf, "r+b") as f:
        f.truncate(11 * fb + 5)
    with pytest.raises(N.NativeError, match="short read"):
        m.predict_video_source(vs)


def test_loss_is_ten_minus_jod_and_refuses_gradients(mock_device):
    """cvvdp.loss (cvvdp_metric.py:294-298): the value, and a clear refusal when a gradient is expected."""
    tst, ref = synth.make_pair_u8(33, 1, 24, 40)
    m = cv.cvvdp(display_name="standard_fhd")
    t, r = torch.from_numpy(tst[0, :, 0]).float() / 255, torch.from_numpy(ref[0, :, 0]).float() / 255
    jod, _ = m.predict(t, r, dim_order="CHW")
    assert float(m.loss(t, r, dim_order="CHW")) == pytest.approx(10.0 - float(jod), abs=1e-6)
    with pytest.raises(NotImplementedError, match="backward"):
        m.loss(t.clone().requires_grad_(True), r, dim_order="CHW")


def test_yuv_filename_metadata():
    p = cv.decode_video_props("/x/clip_1280x720_10b_444_2020_59.94fps.yuv")
    assert (p["width"], p["height"], p["bit_depth"], p["chroma_ss"], p["color_space"], p["fps"]) == (1280, 720, 10, "444", "2020", 59.94)
    assert cv.decode_video_props("a_640x480p30_hdr.yuv")["fps"] == 30
    assert cv.create_yuv_fname("b", p) == "b_1280x720_10b_444_2020_59.94fps.yuv"


@pytest.mark.parametrize("name", gu.feature_case_names())
def test_features_on_mock_device(name, mock_device):
    """cvvdp.extract_features (band kernel in feature mode + k_feature_pool) against tensors
    produced by the reference's cvvdp_ml_base.extract_features, and the ordinary prediction is unchanged
    by switching between the two plans."""
    z, meta = gu.load_case(name)
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"])
    vs = cv.video_source_array(z["test"], z["ref"], meta["fps"], dim_order=meta["dim_order"], display_photometry=m.display_photometry)
    j0, s0 = m.predict_video_source(vs)
    feats, hm = m.extract_features(vs)
    assert hm is None
    gu.assert_features_close([f.numpy() for f in feats], z, name)
    j1, s1 = m.predict_video_source(vs)
    assert np.array_equal(s0["Q_per_ch"], s1["Q_per_ch"]) and torch.equal(torch.as_tensor(j0), torch.as_tensor(j1))


def test_get_temporal_filters_matches_reference(mock_device):
    """cvvdp.get_temporal_filters against the reference's irfft-based filters (cvvdp_metric.py:1057-1092)."""
    import os
    z = np.load(os.path.join(gu.GOLDEN_DIR, "known_answer_temporal_filters.npz"))
    m = cv.cvvdp(display_name="standard_4k")
    for key in z.files:
        if not key.startswith("fps_"):
            continue
        F, omega = m.get_temporal_filters(float(key[4:]))
        got = np.stack([f.numpy() for f in F])
        assert got.shape == z[key].shape
        assert np.max(np.abs(got - z[key])) <= 2e-6, key
        assert np.array_equal(got, got[:, ::-1])  # exactly symmetric
        assert np.array_equal(omega.numpy(), z["omega_bands"])


@pytest.mark.parametrize("fps", [8, 15, 24, 30, 40, 48, 50, 60, 120])
def test_every_temporal_specialisation_against_oracle(fps, mock_device):
    """Filter lengths 3..17 take the two-stage kernel specialised for that length (chunk sizes 2..9, clips
    shorter and longer than the filter, symmetric padding); 31 taps (120 fps) take the generic kernel."""
    F = 7 if fps < 60 else 21
    tst, ref = synth.make_pair_u8(60 + fps, F, 16, 64)  # 1024 pixels: whole 64-pixel warp segments
    m = cv.cvvdp(display_name="standard_fhd", temp_padding="symmetric")
    jod, stats = m.predict(tst, ref, frames_per_second=fps)
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, "standard_fhd", "symmetric")
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"{fps} fps")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


@pytest.mark.parametrize("fps,dtype,display,padding", [(72, "u8", "standard_fhd", "replicate"), (90, "f32", "standard_4k", "symmetric"),
                                                       (120, "u8", "standard_4k", "symmetric"), (120, "f16", "standard_hdr_linear", "replicate"),
                                                       (165, "u16", "standard_hdr_pq", "replicate")])
def test_long_filters_take_the_shared_ring_kernel(fps, dtype, display, padding, mock_device):
    """Frame rates above 64 fps (19, 25, 31, 43 taps): packed shared-memory-ring temporal kernel, table and float
    variants, clips shorter and longer than the filter, against the oracle."""
    F = 12 if fps == 90 else 50
    tst, ref = synth.make_pair_u8(70 + fps, F, 16, 64)
    if dtype == "u8":
        tst_in, ref_in = tst, ref
    elif dtype == "u16":
        tst_in, ref_in = tst.astype(np.uint16) * 180, ref.astype(np.uint16) * 180
    else:
        scale = 3.0 if display == "standard_hdr_linear" else 1.0
        tst_in, ref_in = (tst.astype(np.float32) / 255 * scale).astype(dtype.replace("f", "float")), (ref.astype(np.float32) / 255 * scale).astype(dtype.replace("f", "float"))
    jod_o, stats_o = O.predict(tst_in, ref_in, "BCFHW", fps, display, padding)
    if dtype == "u16":
        tst_in, ref_in = tst_in.view(np.int16), ref_in.view(np.int16)
    m = cv.cvvdp(display_name=display, temp_padding=padding)
    jod, stats = m.predict(tst_in, ref_in, frames_per_second=fps)
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"{fps} fps {dtype}")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_image_batches_take_the_one_tap_two_stage_kernel(dtype, mock_device):
    """Images whose planes are whole 64-pixel segments (planar BCHW and interleaved BHWC batches) use the staged,
    table-driven front end with a one-tap FIR; other sizes the generic kernel.  All against the oracle, and the two
    layouts bit-identical."""
    B, H, W = 3, 24, 64
    tst = np.concatenate([synth.make_pair_u8(150 + b, 1, H, W)[0] for b in range(B)], 0)  # [B,3,1,H,W]
    ref = np.concatenate([synth.make_pair_u8(150 + b, 1, H, W)[1] for b in range(B)], 0)
    if dtype == "f32":
        tst, ref = tst.astype(np.float32) / 255, ref.astype(np.float32) / 255
    m = cv.cvvdp(display_name="standard_fhd")
    jod, planar = m.predict(tst[:, :, 0], ref[:, :, 0], dim_order="BCHW")
    jod_o, want = O.predict(tst, ref, "BCFHW", 0, "standard_fhd")
    gu.assert_q_close(planar["Q_per_ch"], want["Q_per_ch"], "image batch")
    assert np.max(np.abs(np.asarray(jod.cpu() if hasattr(jod, "cpu") else jod, dtype=np.float64) - np.asarray(jod_o))) <= gu.JOD_TOL
    ti, ri = np.ascontiguousarray(tst[:, :, 0].transpose(0, 2, 3, 1)), np.ascontiguousarray(ref[:, :, 0].transpose(0, 2, 3, 1))
    _, inter = m.predict(ti, ri, dim_order="BHWC")
    assert np.array_equal(inter["Q_per_ch"], planar["Q_per_ch"])


@pytest.mark.parametrize("fps,dtype", [(30, "u8"), (60, "f32"), (24, "f16"), (120, "u8"), (90, "f32")])
def test_channel_interleaved_frames_take_the_fast_temporal_kernels(fps, dtype, mock_device):
    """FHWC clips (what decoded frames stacked in numpy look like): the two-stage and shared-ring kernels read the
    interleaved pixels in place and give exactly the bits of the planar BCFHW layout; host and device residency."""
    F, H, W = (11, 16, 64) if fps < 64 else (36, 16, 64)
    tst, ref = synth.make_pair_u8(140 + fps, F, H, W)
    if dtype != "u8":
        tst, ref = (tst.astype(np.float32) / 255).astype(dtype.replace("f", "float")), (ref.astype(np.float32) / 255).astype(dtype.replace("f", "float"))
    m = cv.cvvdp(display_name="standard_fhd")
    _, planar = m.predict(tst, ref, frames_per_second=fps)
    ti, ri = np.ascontiguousarray(tst[0].transpose(1, 2, 3, 0)), np.ascontiguousarray(ref[0].transpose(1, 2, 3, 0))  # [F,H,W,C]
    _, host = m.predict(ti, ri, dim_order="FHWC", frames_per_second=fps)
    assert np.array_equal(host["Q_per_ch"], planar["Q_per_ch"])
    vs = cv.video_source_array(ti, ri, fps, dim_order='FHWC', display_photometry=m.display_photometry); Qd, _ = m.q_per_ch_from_tensors(vs.test_video, vs.reference_video, F, fps, _resident=True); dev_res = dict(Q_per_ch=Qd.numpy())
    assert np.array_equal(dev_res["Q_per_ch"], planar["Q_per_ch"])


def test_shared_ring_kernel_is_independent_of_the_frame_partition(mock_device):
    """The shared-ring kernel emits two frames per pass where it can; any split of the clip into frame ranges (odd and
    even lengths, single frames) must give the bits of the whole clip."""
    tst, ref = synth.make_pair_u8(77, 40, 16, 64)
    m = cv.cvvdp(display_name="standard_fhd")
    _, whole = m.predict(tst, ref, frames_per_second=120)
    vs = cv.video_source_array(tst, ref, 120, display_photometry=m.display_photometry)
    parts = np.zeros_like(whole["Q_per_ch"])
    for lo, hi in ((0, 7), (7, 8), (8, 31), (31, 40)):
        _, s = m.predict_video_source(vs, frame_range=(lo, hi))
        parts[:, :, lo:hi] = s["Q_per_ch"][:, :, lo:hi]
    assert np.array_equal(parts, whole["Q_per_ch"])


@pytest.mark.parametrize("dtype,fps", [("f32", 60), ("f16", 30), ("f32", 24)])
def test_two_stage_temporal_kernel_float_inputs(dtype, fps, mock_device):
    """fp32 / fp16 clips with whole 64-pixel warp segments: the non-table variant of the two-stage temporal
    kernel (per-pixel EOTF in the rolled front end, 2- and 4-byte raw stage, more pieces than lanes)."""
    tst, ref = synth.make_pair_u8(80 + fps, 10, 16, 128)
    tst, ref = tst.astype(np.float32) / 255, ref.astype(np.float32) / 255
    if dtype == "f16":
        tst, ref = tst.astype(np.float16), ref.astype(np.float16)
    m = cv.cvvdp(display_name="standard_4k")
    jod, stats = m.predict(tst, ref, frames_per_second=fps)
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, "standard_4k")
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"{dtype} {fps}")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


@pytest.mark.parametrize("F,H,W,fps", [(1, 8, 8, 0), (1, 7, 40, 0), (1, 6, 64, 0), (3, 13, 17, 30), (1, 4, 4, 0),
                                       (2, 5, 300, 24), (1, 300, 5, 0)])
def test_tiny_and_degenerate_sizes(F, H, W, fps, mock_device):
    """Two- and three-band pyramids, levels too small for the phase-uncertainty blur (h or w <= 6,
    cvvdp_metric.py:965), one-strip / one-segment grids, extreme aspect ratios.  The reference accepts all
    of these (checked in the build container: oracle == reference within 1 % of the gate)."""
    tst, ref = synth.make_pair_u8(90 + H + W, F, H, W)
    m = cv.cvvdp(display_name="standard_fhd")
    jod, stats = m.predict(tst, ref, frames_per_second=fps)
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, "standard_fhd")
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"{F}x{H}x{W}")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


def test_pageable_host_clips_take_the_bounce_buffers(mock_device, monkeypatch):
    """Pageable host memory is uploaded through pinned 32 MiB bounce slots filled by several host threads; the
    result must not change by a bit (CVVDP_B200_FORCE_STAGING makes the mock device take that path)."""
    tst, ref = synth.make_pair_u8(95, 24, 36, 64)
    m = cv.cvvdp(display_name="standard_fhd")
    _, direct = m.predict(tst, ref, frames_per_second=60)
    monkeypatch.setenv("CVVDP_B200_FORCE_STAGING", "pool")  # (and parts small enough for the helper threads to share)
    _, staged = m.predict(tst, ref, frames_per_second=60)
    assert np.array_equal(direct["Q_per_ch"], staged["Q_per_ch"])
    for _ in range(3):  # the pool is reused from call to call
        _, again = m.predict(tst, ref, frames_per_second=60)
        assert np.array_equal(direct["Q_per_ch"], again["Q_per_ch"])


def test_mixed_dtypes_and_single_frame_slices(mock_device):
    """Test and reference of different dtypes are accepted like in the reference (it unpacks per frame); pooling a
    one-frame slice of a video keeps the video formula (four channels, no image_int)."""
    tst, ref = synth.make_pair_u8(97, 5, 24, 40)
    m = cv.cvvdp(display_name="standard_fhd")
    j_u8, s_u8 = m.predict(tst, ref, frames_per_second=30)
    j_mix, s_mix = m.predict(tst, ref.astype(np.float32) / 255, frames_per_second=30)
    gu.assert_q_close(s_mix["Q_per_ch"], s_u8["Q_per_ch"], "mixed dtypes")
    assert abs(float(j_mix) - float(j_u8)) <= 1e-4
    vs = cv.video_source_array(tst, ref, 30, display_photometry=m.display_photometry)
    j1, s1 = m.predict_video_source(vs, frame_range=(2, 3))
    j1_o = O.do_pooling_and_jods(s_u8["Q_per_ch"][:, :, 2:3], O.Params(), is_image=False)
    assert abs(float(j1) - float(np.asarray(j1_o).ravel()[0])) <= 1e-4


def _prefiltered_source(z, fps):
    class Prefiltered(cv.video_source):  # a third-party source that does its own temporal filtering
        is_temporally_filtered = True

        def get_video_size(self):
            return (z["test4"].shape[3], z["test4"].shape[4], z["test4"].shape[2])

        def get_frames_per_second(self):
            return fps

        def get_test_frame(self, frame, device, colorspace):
            assert colorspace == "DKLd65_trans"
            return torch.from_numpy(z["test4"][:, :, frame:frame + 1]).to(device)

        def get_reference_frame(self, frame, device, colorspace):
            assert colorspace == "DKLd65_trans"
            return torch.from_numpy(z["ref4"][:, :, frame:frame + 1]).to(device)

    return Prefiltered()


def test_prefiltered_video_source(mock_device):
    """is_temporally_filtered sources (cvvdp_metric.py:470-488): four-channel frames bypass the temporal filter."""
    z, meta = gu.load_case("prefilt_vid_f32_7x48x80_fhd")
    m = cv.cvvdp(display_name=meta["display"])
    jod, stats = m.predict_video_source(_prefiltered_source(z, meta["fps"]))
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], "prefiltered")
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL


@pytest.mark.parametrize("shape,fps,heatmap", [((1, 150, 360), 0, "raw"), ((5, 64, 236), 30, None), ((2, 135, 240), 30, None)])
def test_fused_band_and_reduce_kernel(shape, fps, heatmap, mock_device, monkeypatch):
    """CVVDP_B200_FUSED_REDUCE=1: the band kernel of a level computes the next pyramid level itself (no reduce launches
    for the levels with blur) -- against the oracle and against the default pair of kernels (odd sizes: the
    reference's edge rules incl. the row-parity quirk of lpyr_dec.py:206 inside the fused kernel)."""
    F, H, W = shape
    tst, ref = synth.make_pair_u8(12, F, H, W)
    _, plain = cv.cvvdp(display_name="standard_fhd", heatmap=heatmap).predict(tst, ref, frames_per_second=fps)
    monkeypatch.setenv("CVVDP_B200_FUSED_REDUCE", "1")
    m = cv.cvvdp(display_name="standard_fhd", heatmap=heatmap)
    m._ctx.profile_enable(True)
    jod, stats = m.predict(tst, ref, frames_per_second=fps)
    assert sum(1 for p in m._ctx.profile_read() if p["kind"] == "reduce") < stats["Q_per_ch"].shape[3] - 1
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, "standard_fhd", heatmap=heatmap)
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], str(shape))
    gu.assert_q_close(stats["Q_per_ch"], plain["Q_per_ch"], "fused vs plain")
    if heatmap:
        assert np.max(np.abs(stats["heatmap"].float().numpy() - stats_o["heatmap"].astype(np.float32))) <= gu.HEATMAP_ATOL
