"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path through the C ABI against the
reference-generated fixtures (tests/golden) and the numpy oracle on seeded inputs, plus
size-independent properties at full BASELINE sizes.  Tolerances: |dJOD| <= 1e-3,
|dQ_per_ch| <= 1e-3 |Q| + 1e-5, raw heat map (fp16, 0..1) <= 2e-3."""
import os

import numpy as np
import pytest
import torch

import golden_util as gu
import synth
from oracle import cvvdp_oracle as O

import colorvideovdp_b200 as cv

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _t(a):
    if isinstance(a, np.ndarray) and a.dtype == np.uint16:
        a = a.view(np.int16)
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("resident", ["device", "host"])
@pytest.mark.parametrize("name", gu.case_names())
def test_golden_fixtures(name, resident):
    z, meta = gu.load_case(name)
    # hm_block == 1: the reference coloured this clip frame by frame (its CPU block size); one frame per pass here too
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], heatmap=meta["heatmap"], device=DEV,
                 gpu_mem=1e-6 if meta.get("hm_block") == 1 else None)
    tst, ref = (z["test"], z["ref"]) if resident == "host" else (_t(z["test"]), _t(z["ref"]))
    jod, stats = m.predict(tst, ref, dim_order=meta["dim_order"], frames_per_second=meta["fps"])
    assert jod.device.type == "cuda"
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert np.max(np.abs(jod.cpu().numpy().astype(np.float64) - z["jod"])) <= gu.JOD_TOL
    if meta["heatmap"] == "raw":
        hm = stats["heatmap"]
        assert hm.dtype == torch.float16 and hm.device.type == "cpu"
        assert np.max(np.abs(hm.float().numpy() - z["heatmap"].astype(np.float32))) <= gu.HEATMAP_ATOL
    elif meta["heatmap"] in ("threshold", "supra-threshold"):
        hm = stats["heatmap"]
        assert tuple(hm.shape) == z["heatmap"].shape
        assert np.max(np.abs(hm.float().numpy() - z["heatmap"].astype(np.float32))) <= gu.COLOR_HEATMAP_ATOL
    assert m._ctx.launch_count() > 0


CASES = [  # (seed, F, H, W, fps, display, padding, dtype)
    (31, 1, 256, 256, 0, "standard_fhd", "replicate", "u8"),       # BASELINE config 1
    (32, 12, 135, 240, 30, "standard_fhd", "replicate", "u8"),     # odd H / even W: lpyr_dec.py:206 quirk
    (33, 6, 136, 241, 60, "standard_4k", "symmetric", "f32"),      # even H / odd W, fl=17 > F
    (34, 20, 100, 180, 60, "standard_hdr_pq", "replicate", "u16"),
    (35, 5, 270, 480, 120, "standard_4k", "replicate", "f16"),     # fl=31
    (36, 3, 97, 33, 25, "standard_phone", "symmetric", "u8"),      # tiles with ragged edges, tall image
    (37, 20, 64, 128, 60, "standard_4k", "replicate", "f32"),      # whole 64-pixel segments: two-stage temporal kernel, fp32 input
    (38, 12, 48, 128, 30, "standard_fhd", "symmetric", "f16"),     # two-stage temporal kernel, fp16 input, 9 taps
]


@pytest.mark.parametrize("seed,F,H,W,fps,display,padding,dtype", CASES)
def test_against_oracle(seed, F, H, W, fps, display, padding, dtype):
    if dtype == "u16":
        tst, ref = synth.make_pair_pq_u16(seed, F, H, W)
    else:
        tst, ref = synth.make_pair_u8(seed, F, H, W)
        if dtype == "f32":
            tst, ref = tst.astype(np.float32) / 255, ref.astype(np.float32) / 255
        if dtype == "f16":
            tst, ref = (tst.astype(np.float32) / 255).astype(np.float16), (ref.astype(np.float32) / 255).astype(np.float16)
    m = cv.cvvdp(display_name=display, temp_padding=padding, device=DEV)
    jod, stats = m.predict(_t(tst), _t(ref), frames_per_second=fps)
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, display, padding)
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"seed {seed}")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


def test_heatmap_against_oracle():
    tst, ref = synth.make_pair_u8(41, 4, 120, 200)
    m = cv.cvvdp(display_name="standard_4k", heatmap="raw", device=DEV)
    jod, stats = m.predict(_t(tst), _t(ref), frames_per_second=30)
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", 30, "standard_4k", heatmap="raw")
    assert np.max(np.abs(stats["heatmap"].float().numpy() - stats_o["heatmap"].astype(np.float32))) <= gu.HEATMAP_ATOL
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"])


def test_identical_pair_is_exactly_10_at_full_hd():
    _, ref = synth.make_pair_u8(42, 4, 1080, 1920)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    r = _t(ref)
    jod, stats = m.predict(r, r, frames_per_second=30)
    assert float(jod) == 10.0 and np.all(stats["Q_per_ch"] == 0)


def test_partition_independence_full_hd():
    """Frame blocks, frame shards and host/device residency give bit-identical Q_per_ch at 1080p."""
    tst, ref = synth.make_pair_u8(43, 12, 1080, 1920)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    t, r = _t(tst), _t(ref)
    _, full = m.predict(t, r, frames_per_second=30)
    small = cv.cvvdp(display_name="standard_fhd", device=DEV, gpu_mem=0.5)
    _, blk = small.predict(t, r, frames_per_second=30)
    assert small._info.block_frames < m._info.block_frames
    assert np.array_equal(full["Q_per_ch"], blk["Q_per_ch"])
    _, host = m.predict(tst, ref, frames_per_second=30)
    assert np.array_equal(full["Q_per_ch"], host["Q_per_ch"])
    vs = cv.video_source_array(t, r, 30, display_photometry=m.display_photometry)
    Qa, _ = m.compute_q_per_ch(vs, (0, 5))
    Qb, _ = m.compute_q_per_ch(vs, (5, 12))
    assert np.array_equal((Qa + Qb).cpu().numpy(), full["Q_per_ch"])
    # monotonicity: more noise -> lower JOD
    tst2 = np.clip(tst.astype(np.int16) + (np.random.default_rng(0).integers(-12, 13, tst.shape)), 0, 255).astype(np.uint8)
    j2, _ = m.predict(_t(tst2), r, frames_per_second=30)
    j1 = m.do_pooling_and_jods(full["Q_per_ch"])
    assert float(j2) < float(j1) < 10.0


def test_4k_block_matches_oracle_on_one_frame():
    """BASELINE config 3 shape (3840x2160, 60 fps, standard_4k): the oracle is too slow for the clip, so
    one late frame is checked (needs 17 frames of history)."""
    tst, ref = synth.make_pair_u8(44, 18, 2160, 3840)
    m = cv.cvvdp(display_name="standard_4k", device=DEV)
    vs = cv.video_source_array(_t(tst), _t(ref), 60, display_photometry=m.display_photometry)
    Q, _ = m.compute_q_per_ch(vs, (17, 18))
    _, so = O.predict(tst, ref, "BCFHW", 60, "standard_4k", frame_range=(17, 18))
    gu.assert_q_close(Q.cpu().numpy()[:, :, 17:18], so["Q_per_ch"][:, :, 17:18], "4k frame 17")


def test_display_model_plugin_surface():
    rng = np.random.default_rng(2)
    V = rng.random((1, 3, 1, 64, 80), dtype=np.float32)
    for name in ("standard_4k", "standard_hdr_pq", "standard_hdr_linear", "standard_hdr_hlg"):
        dm = cv.vvdp_display_photometry.load(name, [])
        odm = O.Display(name)
        L = dm.forward(torch.from_numpy(V).to(DEV)).cpu().numpy()
        L_ref = O.eotf_forward(V[:, :, 0], odm)
        assert np.max(np.abs(L[:, :, 0] - L_ref) / np.abs(L_ref)) < 1e-4, name
        D = dm.source_2_target_colorspace(torch.from_numpy(V).to(DEV), "DKLd65").cpu().numpy()
        D_ref = O.frontend(V[:, :, 0], odm)
        assert np.max(np.abs(D[:, :, 0] - D_ref)) < 1e-4 * np.abs(D_ref).max(), name
    vs = cv.video_source_array(V, V, 0, display_photometry="standard_4k")
    fr = vs.get_test_frame(0, DEV, "DKLd65")
    assert tuple(fr.shape) == (1, 3, 1, 64, 80) and fr.dtype == torch.float32 and fr.is_cuda


def test_plugin_source_and_pooling_entry():
    tst, ref = synth.make_pair_u8(45, 11, 90, 120)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    jod, fast = m.predict(_t(tst), _t(ref), frames_per_second=30)

    class Wrapped(cv.video_source):  # a third-party source built on the display-model plugin surface
        def __init__(self):
            self.inner = cv.video_source_array(tst, ref, 30, display_photometry="standard_fhd")

        def get_video_size(self):
            return self.inner.get_video_size()

        def get_frames_per_second(self):
            return 30

        def get_test_frame(self, f, device, colorspace):
            return self.inner.get_test_frame(f, device, colorspace)

        def get_reference_frame(self, f, device, colorspace):
            return self.inner.get_reference_frame(f, device, colorspace)

    jod_p, plug = m.predict_video_source(Wrapped())
    gu.assert_q_close(plug["Q_per_ch"], fast["Q_per_ch"], "plugin vs fused front end")
    assert abs(float(jod_p) - float(jod)) <= 1e-4
    P = O.Params()
    assert abs(float(m.do_pooling_and_jods(fast["Q_per_ch"])) - float(O.do_pooling_and_jods(fast["Q_per_ch"], P))) < 2e-5


def test_config4_hdr_pq_4k_heatmap_one_frame():
    """BASELINE config 4 shape: 3840x2160 HDR (PQ) at 60 fps on standard_hdr_pq with the raw heat map.
    The oracle checks Q_per_ch and the heat map of one late frame (17 frames of history)."""
    tst, ref = synth.make_pair_pq_u16(46, 18, 2160, 3840)
    m = cv.cvvdp(display_name="standard_hdr_pq", heatmap="raw", device=DEV)
    vs = cv.video_source_array(_t(tst), _t(ref), 60, display_photometry=m.display_photometry)
    Q, hm = m.compute_q_per_ch(vs, (17, 18))
    _, so = O.predict(tst, ref, "BCFHW", 60, "standard_hdr_pq", heatmap="raw", frame_range=(17, 18))
    gu.assert_q_close(Q.cpu().numpy()[:, :, 17:18], so["Q_per_ch"][:, :, 17:18], "4k hdr frame 17")
    err = np.abs(hm[0, 0, 17].float().cpu().numpy() - so["heatmap"][0, 0, 17].astype(np.float32))
    assert err.max() <= gu.HEATMAP_ATOL
    # whole 18-frame clip through the public API, host-resident input, heat map returned on the CPU in fp16
    jod, stats = m.predict(tst, ref, frames_per_second=60)
    assert stats["heatmap"].shape == (1, 1, 18, 2160, 3840) and stats["heatmap"].dtype == torch.float16
    assert np.array_equal(stats["Q_per_ch"][:, :, 17], Q.cpu().numpy()[:, :, 17])
    assert 0.0 < float(jod) < 10.0


@pytest.mark.parametrize("name", gu.yuv_case_names())
def test_yuv_files(name, tmp_path):
    """Raw planar YUV ingestion fused into the temporal front end, against the reference fixture."""
    tf, rf, z, meta = gu.write_yuv_case(name, str(tmp_path))
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], device=DEV)
    vs = cv.video_source_yuv_file(tf, rf, display_photometry=meta["display"])
    jod, stats = m.predict_video_source(vs)
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL
    rgb = vs.test_vidr.get_frame_rgb_tensor(z["Q_per_ch"].shape[2] - 1, DEV)
    assert np.max(np.abs(rgb.cpu().numpy() - z["rgb_last_test_frame"])) <= 2e-6


@pytest.mark.parametrize("name", gu.vfile_case_names())
def test_video_file_sources(name, tmp_path, monkeypatch):
    """ffmpeg-pipe readers and the full-screen resize on hardware, against the reference fixtures (the pipe is served by
    tests/fake_ffmpeg: there is no ffmpeg in the image)."""
    from test_emu_parity import open_vfile_source
    monkeypatch.setenv("PATH", gu.FAKE_FFMPEG_DIR + os.pathsep + os.environ["PATH"])
    tf, rf, z, meta = gu.write_vfile_case(name, str(tmp_path))
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], device=DEV)
    vs = open_vfile_source(tf, rf, meta)
    jod, stats = m.predict_video_source(vs)
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], name)
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL
    H, W, F = vs.get_video_size()
    if meta["kind"] == "yuv":
        rd = cv.video_reader_yuv(tf, resize_fn=meta["full_screen_resize"], resize_height=H, resize_width=W)
    else:
        vs2 = open_vfile_source(tf, rf, meta)
        vs2.init_readers()
        rd = vs2.test_vidr
    rgb = rd.unpack(rd.get_frame(), DEV)
    assert np.max(np.abs(rgb.cpu().numpy() - z["rgb_first_test_frame"])) <= 5e-6
    rd.close()


@pytest.mark.parametrize("mode", ["nearest", "bilinear", "bicubic", "area"])
def test_resize_kernel_matches_torch_interpolate(mode):
    """k_resize on hardware against torch.nn.functional.interpolate on the same device, 1080p -> 4K and 4K -> 1080p."""
    from colorvideovdp_b200 import _native as N
    from colorvideovdp_b200 import cvvdp_metric as cm
    params, lut = cm._default_native_inputs()
    ctx = N.Context(params, lut, DEV.index or 0)
    g = torch.Generator().manual_seed(3)
    st = torch.cuda.current_stream(DEV).cuda_stream
    for (H, W, OH, OW) in [(1080, 1920, 2160, 3840), (2160, 3840, 1080, 1920), (270, 480, 777, 1001)]:
        src = (torch.rand((3, H, W), generator=g) * 1.2 - 0.1).to(DEV)
        dst = torch.empty((3, OH, OW), device=DEV)
        ctx.resize(src.data_ptr(), dst.data_ptr(), 3, H, W, OH, OW, mode, True, st)
        want = torch.nn.functional.interpolate(src[None], size=(OH, OW), mode=mode)[0].clip(0, 1)
        assert float((dst - want).abs().max()) <= 5e-6, (mode, H, W, OH, OW)


def test_yuv_4k_file_streams_in_windows(tmp_path):
    """A 4K 4:2:0 10-bit .yuv pair walked in several host windows (the mapping is read by the upload threads of the
    library, no intermediate copy) == one window."""
    from golden.make_golden_yuv_synth import synth_yuv
    F, H, W = 12, 2160, 3840
    t, r = synth_yuv(72, F, H, W, "420", 10)
    props = {"width": W, "height": H, "fps": 30, "bit_depth": 10, "color_space": "2020", "chroma_ss": "420"}
    tf, rf = str(tmp_path / cv.create_yuv_fname("t", props)), str(tmp_path / cv.create_yuv_fname("r", props))
    t.tofile(tf), r.tofile(rf)
    m = cv.cvvdp(display_name="standard_hdr_pq", device=DEV)
    jod, whole = m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry="standard_hdr_pq"))
    m.yuv_chunk_bytes = 1
    jod2, parts = m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry="standard_hdr_pq"))
    assert np.array_equal(parts["Q_per_ch"], whole["Q_per_ch"]) and float(jod) == float(jod2)
    T, _ = O.read_yuv_rgb(tf, frames=F)
    R, _ = O.read_yuv_rgb(rf, frames=F)
    _, rgb_path = m.predict(torch.from_numpy(T).to(DEV), torch.from_numpy(R).to(DEV), frames_per_second=30)
    gu.assert_q_close(whole["Q_per_ch"], rgb_path["Q_per_ch"], "4K yuv vs rgb path")


@pytest.mark.parametrize("chroma,bit_depth,color_space,display", [("420", 8, "709", "standard_fhd"), ("422", 10, "2020", "standard_hdr_pq"),
                                                               ("444", 8, "709", "standard_hdr_hlg"), ("420", 10, "2020", "standard_hdr_pq")])
def test_yuv_two_stage_front_end(chroma, bit_depth, color_space, display, tmp_path, monkeypatch):
    """Planar-YUV front end of the packed two-stage temporal kernel (rows of whole 64-pixel segments) against the oracle
    and the frame-by-frame path."""
    from golden.make_golden_yuv_synth import synth_yuv
    F, H, W = 6, 34, 128
    t, r = synth_yuv(91, F, H, W, chroma, bit_depth)
    props = {"width": W, "height": H, "fps": 30, "bit_depth": bit_depth, "color_space": color_space, "chroma_ss": chroma}
    tf, rf = str(tmp_path / cv.create_yuv_fname("t", props)), str(tmp_path / cv.create_yuv_fname("r", props))
    t.tofile(tf), r.tofile(rf)
    m = cv.cvvdp(display_name=display, device=DEV)
    jod, fast = m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry=display))
    jod_o, want = O.predict_yuv(tf, rf, display)
    gu.assert_q_close(fast["Q_per_ch"], want["Q_per_ch"], "two-stage YUV vs oracle")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL
    vs = cv.video_source_yuv_file(tf, rf, display_photometry=display)
    monkeypatch.setattr(vs, "yuv_readers", lambda: None)
    _, slow = m.predict_video_source(vs)
    gu.assert_q_close(fast["Q_per_ch"], slow["Q_per_ch"], "two-stage YUV vs frame by frame")


def test_yuv_1080p_matches_oracle_and_rgb_path(tmp_path):
    """1080p 4:2:0 8-bit clip: fused YUV path == (oracle RGB conversion -> fp32 RGB path)."""
    from golden.make_golden_yuv_synth import synth_yuv  # shared synthetic YUV generator
    F, H, W = 10, 1080, 1920
    t, r = synth_yuv(71, F, H, W, "420", 8)
    props = {"width": W, "height": H, "fps": 30, "bit_depth": 8, "color_space": "709", "chroma_ss": "420"}
    tf, rf = str(tmp_path / cv.create_yuv_fname("t", props)), str(tmp_path / cv.create_yuv_fname("r", props))
    t.tofile(tf), r.tofile(rf)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    jod, stats = m.predict_video_source(cv.video_source_yuv_file(tf, rf, display_photometry="standard_fhd"))
    T, _ = O.read_yuv_rgb(tf)
    R, _ = O.read_yuv_rgb(rf)
    jod2, stats2 = m.predict(torch.from_numpy(T).to(DEV), torch.from_numpy(R).to(DEV), frames_per_second=30)
    gu.assert_q_close(stats["Q_per_ch"], stats2["Q_per_ch"], "yuv vs rgb path")
    assert abs(float(jod) - float(jod2)) <= 1e-4


@pytest.mark.parametrize("name", gu.feature_case_names())
def test_features_against_reference_fixtures(name):
    """extract_features (band kernel in feature mode + k_feature_pool) against the tensors the
    reference's cvvdp_ml_base.extract_features produced, and an ordinary prediction after it is unchanged."""
    z, meta = gu.load_case(name)
    m = cv.cvvdp(display_name=meta["display"], temp_padding=meta["padding"], device=DEV)
    vs = cv.video_source_array(_t(z["test"]), _t(z["ref"]), meta["fps"], dim_order=meta["dim_order"],
                               display_photometry=m.display_photometry)
    j0, s0 = m.predict_video_source(vs)
    feats, hm = m.extract_features(vs)
    assert hm is None and all(f.device.type == "cuda" for f in feats)
    gu.assert_features_close([f.cpu().numpy() for f in feats], z, name)
    j1, s1 = m.predict_video_source(vs)
    assert np.array_equal(s0["Q_per_ch"], s1["Q_per_ch"]) and torch.equal(j0, j1)


def test_features_1080p_against_oracle():
    """Feature mode at a BASELINE size (1080p, 38-pixel patches, ragged last patch row) against the oracle."""
    tst, ref = synth.make_pair_u8(41, 2, 1080, 1920)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    vs = cv.video_source_array(_t(tst), _t(ref), 30, dim_order="BCFHW", display_photometry=m.display_photometry)
    feats, _ = m.extract_features(vs)
    _, so = O.predict(tst, ref, "BCFHW", 30, "standard_fhd", features=True)
    z = {f"features_b{bb}": f for bb, f in enumerate(so["features"])}

    class Z(dict):
        files = list(z)

    gu.assert_features_close([f.cpu().numpy() for f in feats], Z(z), "1080p")


@pytest.mark.parametrize("fps", [8, 15, 24, 40, 48, 50, 120])
def test_every_temporal_specialisation_on_hardware(fps):
    """Filter lengths 3, 5, 7, 11, 13, 15 (two-stage kernel, one specialisation each) and 31 (generic kernel)
    against the oracle; 9 and 17 taps are covered by the 30 / 60 fps cases above."""
    F = 7 if fps < 60 else 21
    tst, ref = synth.make_pair_u8(60 + fps, F, 32, 64)
    m = cv.cvvdp(display_name="standard_fhd", temp_padding="symmetric", device=DEV)
    jod, stats = m.predict(_t(tst), _t(ref), frames_per_second=fps)
    jod_o, stats_o = O.predict(tst, ref, "BCFHW", fps, "standard_fhd", "symmetric")
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"{fps} fps")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


@pytest.mark.parametrize("fps,dtype,display,padding", [(72, "u8", "standard_fhd", "replicate"), (90, "f32", "standard_4k", "symmetric"),
                                                       (120, "u8", "standard_4k", "symmetric"), (120, "f16", "standard_hdr_linear", "replicate"),
                                                       (165, "u16", "standard_hdr_pq", "replicate")])
def test_long_filters_take_the_shared_ring_kernel(fps, dtype, display, padding):
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
    tst_in, ref_in = _t(tst_in), _t(ref_in)
    m = cv.cvvdp(display_name=display, temp_padding=padding, device=DEV)
    jod, stats = m.predict(tst_in, ref_in, frames_per_second=fps)
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], f"{fps} fps {dtype}")
    assert abs(float(jod) - float(jod_o)) <= gu.JOD_TOL


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_image_batches_take_the_one_tap_two_stage_kernel(dtype):
    """Images whose planes are whole 64-pixel segments (planar BCHW and interleaved BHWC batches) use the staged,
    table-driven front end with a one-tap FIR; other sizes the generic kernel.  All against the oracle, and the two
    layouts bit-identical."""
    B, H, W = 3, 24, 64
    tst = np.concatenate([synth.make_pair_u8(150 + b, 1, H, W)[0] for b in range(B)], 0)  # [B,3,1,H,W]
    ref = np.concatenate([synth.make_pair_u8(150 + b, 1, H, W)[1] for b in range(B)], 0)
    if dtype == "f32":
        tst, ref = tst.astype(np.float32) / 255, ref.astype(np.float32) / 255
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    jod, planar = m.predict(tst[:, :, 0], ref[:, :, 0], dim_order="BCHW")
    jod_o, want = O.predict(tst, ref, "BCFHW", 0, "standard_fhd")
    gu.assert_q_close(planar["Q_per_ch"], want["Q_per_ch"], "image batch")
    assert np.max(np.abs(np.asarray(jod.cpu() if hasattr(jod, "cpu") else jod, dtype=np.float64) - np.asarray(jod_o))) <= gu.JOD_TOL
    ti, ri = np.ascontiguousarray(tst[:, :, 0].transpose(0, 2, 3, 1)), np.ascontiguousarray(ref[:, :, 0].transpose(0, 2, 3, 1))
    _, inter = m.predict(ti, ri, dim_order="BHWC")
    assert np.array_equal(inter["Q_per_ch"], planar["Q_per_ch"])


@pytest.mark.parametrize("fps,dtype", [(30, "u8"), (60, "f32"), (24, "f16"), (120, "u8"), (90, "f32")])
def test_channel_interleaved_frames_take_the_fast_temporal_kernels(fps, dtype):
    """FHWC clips (what decoded frames stacked in numpy look like): the two-stage and shared-ring kernels read the
    interleaved pixels in place and give exactly the bits of the planar BCFHW layout; host and device residency."""
    F, H, W = (11, 16, 64) if fps < 64 else (36, 16, 64)
    tst, ref = synth.make_pair_u8(140 + fps, F, H, W)
    if dtype != "u8":
        tst, ref = (tst.astype(np.float32) / 255).astype(dtype.replace("f", "float")), (ref.astype(np.float32) / 255).astype(dtype.replace("f", "float"))
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    _, planar = m.predict(tst, ref, frames_per_second=fps)
    ti, ri = np.ascontiguousarray(tst[0].transpose(1, 2, 3, 0)), np.ascontiguousarray(ref[0].transpose(1, 2, 3, 0))  # [F,H,W,C]
    _, host = m.predict(ti, ri, dim_order="FHWC", frames_per_second=fps)
    assert np.array_equal(host["Q_per_ch"], planar["Q_per_ch"])
    _, dev_res = m.predict(_t(ti), _t(ri), dim_order='FHWC', frames_per_second=fps)
    assert np.array_equal(dev_res["Q_per_ch"], planar["Q_per_ch"])


@pytest.mark.parametrize("shape", [(6, 16, 64), (3, 20, 28)])  # two-stage temporal kernel / generic kernel
@pytest.mark.parametrize("resident", ["device", "host"])
def test_input_validation_on_the_fused_path(shape, resident, caplog):
    """1.2 -> 'Pixel outside the valid range 0-1' + clamp; NaN -> warning + AssertionError('Must not be nan')
    (display_model.py:335-337, video_source.py:48-59, cvvdp_metric.py:906-907), through predict()."""
    import logging
    F, H, W = shape
    tst, ref = synth.make_pair_u8(21, F, H, W)
    tf, rf = tst.astype(np.float32) / 255, ref.astype(np.float32) / 255
    put = (lambda a: a) if resident == "host" else _t
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    with caplog.at_level(logging.WARNING):
        m.predict(put(tf), put(rf), frames_per_second=30)
    assert not caplog.records
    hot = tf.copy()
    hot[0, 1, 2, 5, 7] = 1.2
    with caplog.at_level(logging.WARNING):
        _, s_hot = m.predict(put(hot), put(rf), frames_per_second=30)
    assert any("Pixel outside the valid range 0-1" in r.message for r in caplog.records)
    hot[0, 1, 2, 5, 7] = 1.0
    _, s_cl = m.predict(put(hot), put(rf), frames_per_second=30)
    assert np.array_equal(s_hot["Q_per_ch"], s_cl["Q_per_ch"])
    caplog.clear()
    bad = tf.copy()
    bad[0, 0, 1, 3, 3] = np.nan
    with caplog.at_level(logging.WARNING):
        with pytest.raises(AssertionError, match="Must not be nan"):
            m.predict(put(bad), put(rf), frames_per_second=30)
    assert any("NaN" in r.message for r in caplog.records)


def test_pageable_numpy_input_equals_pinned_input():
    """predict() on ordinary (pageable) numpy arrays -- uploaded through the library's pinned bounce buffers --
    gives the same bits as pinned tensors and as device-resident tensors (1080p, 24 frames)."""
    tst, ref = synth.make_pair_u8(96, 24, 1080, 1920)
    m = cv.cvvdp(display_name="standard_fhd", device=DEV)
    _, pageable = m.predict(tst, ref, frames_per_second=30)
    tp = torch.from_numpy(tst).pin_memory()
    rp = torch.from_numpy(ref).pin_memory()
    _, pinned = m.predict(tp, rp, frames_per_second=30)
    _, dev = m.predict(_t(tst), _t(ref), frames_per_second=30)
    assert np.array_equal(pageable["Q_per_ch"], pinned["Q_per_ch"]) and np.array_equal(pinned["Q_per_ch"], dev["Q_per_ch"])


def test_prefiltered_video_source_on_hardware():
    """is_temporally_filtered sources (cvvdp_metric.py:470-488) against the reference-generated fixture."""
    from test_emu_parity import _prefiltered_source
    z, meta = gu.load_case("prefilt_vid_f32_7x48x80_fhd")
    m = cv.cvvdp(display_name=meta["display"], device=DEV)
    jod, stats = m.predict_video_source(_prefiltered_source(z, meta["fps"]))
    gu.assert_q_close(stats["Q_per_ch"], z["Q_per_ch"], "prefiltered")
    assert abs(float(jod) - float(z["jod"])) <= gu.JOD_TOL


def test_fused_band_and_reduce_kernel_on_hardware(monkeypatch):
    """CVVDP_B200_FUSED_REDUCE=1 (band kernel computes the next level itself) against the default pair at 1080p and
    against the oracle on an odd-sized clip."""
    tst, ref = synth.make_pair_u8(99, 6, 1080, 1920)
    _, plain = cv.cvvdp(display_name="standard_fhd", device=DEV).predict(_t(tst), _t(ref), frames_per_second=30)
    monkeypatch.setenv("CVVDP_B200_FUSED_REDUCE", "1")
    _, fused = cv.cvvdp(display_name="standard_fhd", device=DEV).predict(_t(tst), _t(ref), frames_per_second=30)
    gu.assert_q_close(fused["Q_per_ch"], plain["Q_per_ch"], "fused vs plain")
    t2, r2 = synth.make_pair_u8(32, 12, 135, 240)
    jod, stats = cv.cvvdp(display_name="standard_fhd", device=DEV).predict(_t(t2), _t(r2), frames_per_second=30)
    jod_o, stats_o = O.predict(t2, r2, "BCFHW", 30, "standard_fhd")
    gu.assert_q_close(stats["Q_per_ch"], stats_o["Q_per_ch"], "fused, odd size")
