"""Golden fixtures for the ffmpeg-pipe readers and the full-screen resize, generated with the UNMODIFIED reference
(pycvvdp.video_source_file.video_source_video_file / pycvvdp.video_source_yuv.video_source_yuv_file +
cvvdp.predict_video_source, CPU).  Container only:

    python tests/golden/make_golden_video_file.py

There is no ffmpeg in the image, so the `ffmpeg` module the reference imports is replaced by a stand-in that "decodes"
a file that already holds raw frames in the pixel format the reference asks for: `probe` returns the stream
description stored next to the file, `run_async` opens the file as the pipe.  Everything after the pipe -- frame
slicing, `unpack`, `_fixed2float_upscale`, the YCbCr matrix, the resize, the display model, the metric -- is the
reference's own code.

Each tests/golden/vfile_*.npz stores the raw bytes of the test/reference streams, the ffprobe stream description, the
constructor arguments and the reference outputs (JOD, Q_per_ch, the RGB tensor `unpack` returned for the first test
frame)."""
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
from make_golden_yuv_synth import synth_yuv  # noqa: E402

pycvvdp = ref_loader.load(prefer_staged=False)
import importlib  # noqa: E402

ref_vsf = importlib.import_module("pycvvdp.video_source_file")  # (the package re-exports a class of the same name)
ref_yuv = importlib.import_module("pycvvdp.video_source_yuv")

DEV = torch.device("cpu")


# ---- stand-in for the ffmpeg-python calls the reference makes (video_source_file.py:80-88,137-148,249-258) ----------
class _Stream:
    def __init__(self, fname):
        self.fname, self.pix_fmt, self.scale = fname, None, None

    def global_args(self, *a):
        return self


class _Process:
    def __init__(self, fname):
        self.stdout = open(fname, "rb")

    def kill(self):
        pass


def _probe(fname, **kw):
    with open(fname + ".probe.json") as f:
        return json.load(f)


def _input(fname):
    return _Stream(fname)


def _filter(stream, name, *args, **kw):
    raise RuntimeError("the stand-in cannot scale: ffmpeg_cc fixtures are generated without resize")


def _output(stream, dst, format=None, pix_fmt=None):
    with open(stream.fname + ".probe.json") as f:
        assert json.load(f)["raw_pix_fmt"] == pix_fmt, f"the file holds a different pixel format than {pix_fmt}"
    return stream


def _run_async(stream, pipe_stdout=True):
    return _Process(stream.fname)


for mod in (sys.modules["ffmpeg"], ref_vsf.ffmpeg):
    mod.probe, mod.input, mod.filter, mod.output, mod.run_async = _probe, _input, _filter, _output, _run_async


def probe_dict(W, H, fps, pix_fmt, raw_pix_fmt, frames, color_space=None, color_transfer=None):
    st = {"codec_type": "video", "width": W, "height": H, "pix_fmt": pix_fmt, "r_frame_rate": f"{fps}/1",
          "nb_frames": str(frames)}
    if color_space:
        st["color_space"] = color_space
    if color_transfer:
        st["color_transfer"] = color_transfer
    return {"streams": [{"codec_type": "audio"}, st], "format": {}, "raw_pix_fmt": raw_pix_fmt}


def run_reference(vs, display, padding):
    m = pycvvdp.cvvdp(display_name=display, device=DEV, temp_padding=padding, quiet=True)
    with torch.no_grad():
        q, s = m.predict_video_source(vs)
    return np.asarray(q.numpy(), dtype=np.float32), s["Q_per_ch"].astype(np.float32)


def save_pipe(name, seed, F, H, W, fps, chroma, bit_depth, color_space, display, padding="replicate", resize=None,
              resize_resolution=None, ffmpeg_cc=False, color_transfer=None):
    """A clip 'decoded by ffmpeg': planar YUV (GPU colour conversion in the reference) or packed RGB (ffmpeg_cc)."""
    if ffmpeg_cc:
        t8, r8 = synth.make_pair_u8(seed, F, H, W)  # [1,3,F,H,W]
        streams = []
        for clip in (t8, r8):
            hwc = np.ascontiguousarray(np.transpose(clip[0], (1, 2, 3, 0)))  # [F,H,W,3]
            streams.append((hwc.astype(np.uint16) * 257) if bit_depth > 8 else hwc)
        t, r = (a.reshape(-1) for a in streams)
        pix_fmt = f"yuv{chroma}p{bit_depth}le" if bit_depth > 8 else f"yuv{chroma}p"
        raw_fmt = "rgb48le" if bit_depth > 8 else "rgb24"
    else:
        t, r = synth_yuv(seed, F, H, W, chroma, bit_depth)
        pix_fmt = raw_fmt = f"yuv{chroma}p{bit_depth}le" if bit_depth > 8 else f"yuv{chroma}p"
    probe = probe_dict(W, H, fps, pix_fmt, raw_fmt, F, color_space, color_transfer)
    with tempfile.TemporaryDirectory() as td:
        tf, rf = os.path.join(td, "test.mp4"), os.path.join(td, "ref.mp4")
        for fn, a in ((tf, t), (rf, r)):
            a.tofile(fn)
            with open(fn + ".probe.json", "w") as f:
                json.dump(probe, f)
        kw = dict(display_photometry=display, full_screen_resize=resize, resize_resolution=resize_resolution,
                  ffmpeg_cc=ffmpeg_cc)
        vs = ref_vsf.video_source_video_file(tf, rf, **kw)
        jod, Q = run_reference(vs, display, padding)
        vs2 = ref_vsf.video_source_video_file(tf, rf, **kw)
        vs2.init_readers()
        rgb = vs2.test_vidr.unpack(vs2.test_vidr.get_frame(), DEV).numpy()
    meta = {"kind": "pipe", "display": display, "padding": padding, "probe": probe, "full_screen_resize": resize,
            "resize_resolution": resize_resolution, "ffmpeg_cc": ffmpeg_cc}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), test_bytes=t, ref_bytes=r, meta=np.asarray(json.dumps(meta)),
                        jod=jod, Q_per_ch=Q, rgb_first_test_frame=rgb.astype(np.float32))
    print(name, float(jod), Q.shape, rgb.shape)


def save_yuv_resized(name, seed, F, H, W, fps, chroma, bit_depth, color_space, display, resize, resize_resolution,
                     retain_aspect_ratio=False, padding="replicate"):
    """video_source_yuv_file with full_screen_resize (video_source_yuv.py:264-338)."""
    t, r = synth_yuv(seed, F, H, W, chroma, bit_depth)
    props = {"width": W, "height": H, "fps": fps, "bit_depth": bit_depth, "color_space": color_space, "chroma_ss": chroma}
    with tempfile.TemporaryDirectory() as td:
        tf = os.path.join(td, ref_yuv.create_yuv_fname("test", props))
        rf = os.path.join(td, ref_yuv.create_yuv_fname("ref", props))
        t.tofile(tf)
        r.tofile(rf)
        vs = ref_yuv.video_source_yuv_file(tf, rf, display_photometry=display, full_screen_resize=resize,
                                           resize_resolution=resize_resolution, retain_aspect_ratio=retain_aspect_ratio)
        jod, Q = run_reference(vs, display, padding)
        size = vs.get_video_size()
        rd = ref_yuv.video_reader_yuv(tf, resize_fn=resize, resize_height=size[0], resize_width=size[1])
        rgb = rd.unpack(rd.get_frame(), DEV).numpy()
        meta = {"kind": "yuv", "test_name": os.path.basename(tf), "ref_name": os.path.basename(rf), "display": display,
                "padding": padding, "full_screen_resize": resize, "resize_resolution": list(resize_resolution),
                "retain_aspect_ratio": retain_aspect_ratio, "video_size": [int(x) for x in size]}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), test_bytes=t, ref_bytes=r, meta=np.asarray(json.dumps(meta)),
                        jod=jod, Q_per_ch=Q, rgb_first_test_frame=rgb.astype(np.float32))
    print(name, float(jod), Q.shape, rgb.shape)


def save_resample(name, seed, F60, H, W, display, padding="replicate"):
    """video_source_temp_resample_file: a 30 fps test stream against a 60 fps reference stream (8-bit 4:2:0)."""
    t, r = synth_yuv(seed, F60, H, W, "420", 8)
    fpix = H * W * 3 // 2
    t30 = np.concatenate([t[f * fpix:(f + 1) * fpix] for f in range(0, F60, 2)])
    probes = (probe_dict(W, H, 30, "yuv420p", "yuv420p", F60 // 2, "bt709"), probe_dict(W, H, 60, "yuv420p", "yuv420p", F60, "bt709"))
    with tempfile.TemporaryDirectory() as td:
        tf, rf = os.path.join(td, "test.mp4"), os.path.join(td, "ref.mp4")
        for fn, a, pr in ((tf, t30, probes[0]), (rf, r, probes[1])):
            a.tofile(fn)
            with open(fn + ".probe.json", "w") as f:
                json.dump(pr, f)
        vs = ref_vsf.video_source_temp_resample_file(tf, rf, display_photometry=display)
        fps, size = vs.get_frames_per_second(), [int(x) for x in vs.get_video_size()]
        jod, Q = run_reference(vs, display, padding)
    meta = {"kind": "resample", "display": display, "padding": padding, "probe_test": probes[0], "probe_ref": probes[1],
            "resample_fps": float(fps), "video_size": size}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), test_bytes=t30, ref_bytes=r, meta=np.asarray(json.dumps(meta)),
                        jod=jod, Q_per_ch=Q)
    print(name, float(jod), Q.shape, fps, size)


if __name__ == "__main__":
    save_resample("vresample_30_vs_60fps_8x40x56", 89, 8, 40, 56, "standard_fhd")
    save_pipe("vfile_pipe_420_8b_bt709_5x48x64", 81, 5, 48, 64, 24, "420", 8, "bt709", "standard_fhd")
    save_pipe("vfile_pipe_422_10b_bt2020nc_4x36x52_pq_bilinear", 82, 4, 36, 52, 30, "422", 10, "bt2020nc", "standard_hdr_pq",
              resize="bilinear", resize_resolution=(78, 54), color_transfer="smpte2084")
    save_pipe("vfile_pipe_444_12b_unknown_12x40x56_sym", 83, 12, 40, 56, 25, "444", 12, None, "standard_4k", padding="symmetric")
    save_pipe("vfile_cc_rgb24_4x40x56", 84, 4, 40, 56, 30, "420", 8, "bt709", "standard_fhd", ffmpeg_cc=True)
    save_pipe("vfile_cc_rgb48_3x36x48", 85, 3, 36, 48, 30, "420", 10, "bt2020nc", "standard_hdr_pq", ffmpeg_cc=True)
    save_yuv_resized("vfile_yuv_420_8b_4x36x48_bicubic", 86, 4, 36, 48, 24, "420", 8, "709", "standard_fhd", "bicubic", (80, 60))
    save_yuv_resized("vfile_yuv_444_10b_3x48x72_area", 87, 3, 48, 72, 30, "444", 10, "2020", "standard_hdr_pq", "area", (45, 30))
    save_yuv_resized("vfile_yuv_422_8b_3x40x64_nearest_aspect", 88, 3, 40, 64, 30, "422", 8, "709", "standard_4k", "nearest",
                     (100, 100), retain_aspect_ratio=True)
