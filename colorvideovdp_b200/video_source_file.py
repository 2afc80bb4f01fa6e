"""Video files decoded by ffmpeg as a video source (pycvvdp/video_source_file.py:74-482).

The reference starts one `ffmpeg` process per clip, reads raw frames from its stdout and converts them with torch
ops (`video_reader_yuv_pytorch.unpack` / `_fixed2float_upscale`, video_source_file.py:261-324) or lets ffmpeg do the
colour conversion (`video_reader`, `--ffmpeg-cc`).  Here the wire format is the same -- `yuv4xxp[NNle]` or
`rgb24` / `rgb48le` raw frames on a pipe -- and the arithmetic is the CUDA front end of this package:

* `cvvdp.predict_video_source` recognises a `video_source_video_file` whose two readers deliver planar YUV of the same
  format and no resize, and streams the raw frames into the fused temporal kernel (`yuv_readers`, the same path as
  raw .yuv files);
* otherwise frames are pulled one by one through `get_test_frame` / `get_reference_frame`: `k_frontend` (unpack,
  chroma upsampling, YCbCr matrix, clip) -> `k_resize` (optional full-screen resize) -> `k_frontend` (display model).

The reference drives ffmpeg through the `ffmpeg-python` package; this module runs the `ffprobe` / `ffmpeg`
executables directly with the arguments that package would generate (`ffprobe -show_format -show_streams -of json`,
`ffmpeg -i FILE -f rawvideo -pix_fmt FMT pipe:`).  Both must be on PATH (or passed as `ffmpeg_cmd` / `ffprobe_cmd`);
without them only .yuv pairs work.  Image files (`video_source_image_frames`, `load_image_as_array`) need imageio /
pyexr and are not part of this package: pass decoded images as arrays to `cvvdp.predict`.
"""
import json
import logging
import math
import os
import re
import shutil
import subprocess

import numpy as np
import torch

from . import _native as N
from .video_source import reshuffle_dims, video_source_dm
from .video_source_yuv import resize_planes, video_reader_yuv, yuv_frame_decoder
from .vq_metric import vq_exception


def ffprobe(vidfile, count_frames=False, cmd="ffprobe"):
    """Stream metadata as the dict `ffmpeg.probe` returns (video_source_file.py:80-88)."""
    exe = shutil.which(cmd)
    if exe is None:
        raise vq_exception(f'ffmpeg failed to open file "{vidfile}" ({cmd} not found)')
    args = [exe, "-show_format", "-show_streams", "-of", "json"] + (["-count_frames"] if count_frames else []) + [vidfile]
    try:
        out = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
        return json.loads(out.decode("utf-8"))
    except (subprocess.CalledProcessError, ValueError, OSError):
        raise vq_exception('ffmpeg failed to open file "' + vidfile + '"')


class video_reader:
    """One clip decoded by ffmpeg to packed RGB (`--ffmpeg-cc`; video_source_file.py:72-198).  Subclasses change the
    pixel format requested from ffmpeg and `unpack`."""

    def __init__(self, vidfile, frames=-1, resize_fn=None, resize_height=-1, resize_width=-1, verbose=False,
                 ffmpeg_cmd="ffmpeg", ffprobe_cmd="ffprobe"):
        if not os.path.isfile(vidfile):
            raise vq_exception('File "' + vidfile + '" not found')
        do_count_frames = vidfile.lower().endswith(".y4m") or frames == -2  # slower but exact
        probe = ffprobe(vidfile, do_count_frames, ffprobe_cmd)
        vstream = next((st for st in probe.get("streams", []) if st.get("codec_type") == "video"), None)
        if vstream is None:
            raise vq_exception('ffmpeg failed to open file "' + vidfile + '"')
        self.fname = vidfile
        self.width = self.src_width = int(vstream["width"])
        self.height = self.src_height = int(vstream["height"])
        self.color_space = vstream.get("color_space", "unknown")
        self.color_transfer = vstream.get("color_transfer", "unknown")
        self.in_pix_fmt = vstream["pix_fmt"]
        num, den = [float(x) for x in vstream["r_frame_rate"].split("/")]
        self.avg_fps = num / den
        if "nb_read_frames" in vstream:
            in_stream = int(vstream["nb_read_frames"])
        elif "nb_frames" in vstream:
            in_stream = int(vstream["nb_frames"])
        elif "DURATION" in vstream.get("tags", {}):  # some VP9 streams carry only a duration
            hrs, mins, secs = map(float, vstream["tags"]["DURATION"].split(":"))
            in_stream = int(np.floor(((hrs * 60 + mins) * 60 + secs) * self.avg_fps))
        else:
            in_stream = -1
        if frames < 0:
            self.frames = in_stream
        else:
            self.frames = frames if in_stream == -1 else min(in_stream, frames)
        self.process = None
        self._ffmpeg_cmd = ffmpeg_cmd
        self._setup_ffmpeg(vidfile, resize_fn, resize_height, resize_width, verbose)
        self.curr_frame = -1

    # ---- ffmpeg process ------------------------------------------------------------------------
    def _start(self, vidfile, out_pix_fmt, verbose, scale=None):
        exe = shutil.which(self._ffmpeg_cmd)
        if exe is None:
            raise vq_exception(f'ffmpeg failed to open file "{vidfile}" ({self._ffmpeg_cmd} not found)')
        args = [exe, "-i", vidfile]
        if scale is not None:
            args += ["-vf", "scale={}:{}:flags={}".format(*scale)]
        args += ["-f", "rawvideo", "-pix_fmt", out_pix_fmt, "pipe:", "-loglevel", "info" if verbose else "quiet"]
        self.process = subprocess.Popen(args, stdout=subprocess.PIPE)

    def _setup_ffmpeg(self, vidfile, resize_fn, resize_height, resize_width, verbose):
        if any(f"p{bd}" in self.in_pix_fmt for bd in (10, 12, 14, 16)):
            out_pix_fmt, self.bpp, self.dtype = "rgb48le", 6, np.uint16
        else:
            out_pix_fmt, self.bpp, self.dtype = "rgb24", 3, np.uint8
        scale = None
        if resize_fn is not None and (resize_width != self.width or resize_height != self.height):
            scale = (resize_width, resize_height, resize_fn if resize_fn != "nearest" else "neighbor")  # ffmpeg resizes
            self.width, self.height = resize_width, resize_height
        self.frame_bytes = int(self.width * self.height * self.bpp)
        self._ctx = None
        self._start(vidfile, out_pix_fmt, verbose, scale)

    def _read_exactly(self, n):
        buf = self.process.stdout.read(n)
        while buf and len(buf) < n:  # a pipe may deliver a frame in pieces
            more = self.process.stdout.read(n - len(buf))
            if not more:
                break
            buf += more
        return buf

    def get_frame(self):
        in_bytes = self._read_exactly(self.frame_bytes) if self.process is not None else b""
        if not in_bytes or len(in_bytes) < self.frame_bytes or (self.frames != -1 and self.curr_frame == self.frames):
            return None
        self.curr_frame += 1
        return np.frombuffer(in_bytes, self.dtype)

    def unpack(self, frame_np, device):
        """Packed RGB samples [H*W*3] -> [H,W,3] fp32 in 0..1 on `device` (video_source_file.py:163-176): the dtype
        unpack of k_frontend on a strided view, pass-through display."""
        ctx, device, stream = yuv_frame_decoder._device_ctx(self, device)
        raw = np.array(frame_np, copy=True).reshape(-1)
        t = torch.from_numpy(raw.view(np.int16) if raw.dtype == np.uint16 else raw).to(device)
        clip = N.Clip()
        clip.data = t.data_ptr()
        clip.stride[1], clip.stride[3], clip.stride[4] = 1, self.width * 3, 3
        clip.frame0, clip.n_frames = 0, 1
        out = torch.empty((3, self.height, self.width), dtype=torch.float32, device=device)
        ctx.frontend(clip, 1, 3, self.height, self.width, N.DTYPE_U16 if raw.dtype == np.uint16 else N.DTYPE_U8, 0,
                     N.CS_RGB_LINEAR, out.data_ptr(), None, stream)
        if stream is not None:
            torch.cuda.current_stream(device).synchronize()
        return out.permute(1, 2, 0)

    def __del__(self):
        self.close()

    def close(self):
        proc = getattr(self, "process", None)
        if proc is not None:
            proc.stdout.close()
            proc.kill()  # ffmpeg would otherwise block on the unread frames
            proc.wait()
            self.process = None

    def __enter__(self):
        return self

    def __exit__(self, type, value, tb):
        self.close()


class video_reader_yuv_pytorch(video_reader, yuv_frame_decoder):
    """ffmpeg only demultiplexes/decodes to planar YUV; unpack, chroma upsampling, colour conversion and the optional
    resize happen on the GPU (video_source_file.py:204-324).  The name is the reference's; the arithmetic is
    k_frontend / k_resize."""

    def _setup_ffmpeg(self, vidfile, resize_fn, resize_height, resize_width, verbose):
        re_grp = re.search(r"p\d+", self.in_pix_fmt)
        self.bit_depth = 8 if re_grp is None else int(re_grp.group().strip("p"))
        self.chroma_ss = self.in_pix_fmt[3:6]
        if self.chroma_ss not in ("444", "420", "422"):
            raise vq_exception(f"GPU-accelerated decoding cannot handle chroma subsampling {self.chroma_ss}. "
                               "Run with `--ffmpeg-cc` command-line argument.")
        out_pix_fmt = f"yuv{self.chroma_ss}p{self.bit_depth}le" if self.bit_depth > 8 else f"yuv{self.chroma_ss}p"
        self.ycbcr_row = "2020" if self.color_space == "bt2020nc" else "bt709"  # video_source_file.py:268-277
        self._init_geometry()
        self.resize_fn = None
        if resize_fn is not None:  # resized later, on the GPU
            if resize_fn not in N.RESIZE_MODES:
                raise vq_exception(f"unknown full-screen resize '{resize_fn}'")
            self.resize_fn, self.resize_height, self.resize_width = resize_fn, resize_height, resize_width
        self._window = []  # (frame index, samples) of the frames_window cache
        self._start(vidfile, out_pix_fmt, verbose)

    def unpack(self, frame_np, device):
        rs = None if self.resize_fn is None else (self.resize_fn, self.resize_height, self.resize_width)
        return self.decode_frame(frame_np, device, resize=rs)

    def frames_window(self, first, count):
        """Frames [first, first+count) as one contiguous host array, reading the pipe forward as needed.  Windows must
        move forward: frames before `first` are dropped."""
        self._window = [(i, a) for (i, a) in self._window if i >= first]
        have = self._window[0][0] if self._window else self.curr_frame + 1
        if have > first:
            raise vq_exception("Video can be currently only read frame-by-frame. Random access not implemented.")
        while self.curr_frame + 1 < first + count:
            a = self.get_frame()
            if a is None:
                raise vq_exception(f'Could not read frame {self.curr_frame + 1} of "{self.fname}". Try passing '
                                   '"--count-frames" or "-nframes".')
            if self.curr_frame >= first:
                self._window.append((self.curr_frame, a))
        out = np.empty(count * self.frame_pixels, dtype=self.dtype)
        for i, a in self._window:
            if first <= i < first + count:
                out[(i - first) * self.frame_pixels:(i - first + 1) * self.frame_pixels] = a
        return out


class video_source_video_file(video_source_dm):
    """Test/reference pair of video files (video_source_file.py:338-482).  The readers start on first use so that the
    object can be pickled before that."""

    def __init__(self, test_fname, reference_fname, display_photometry="sdr_4k_30", config_paths=[], fps=None, frames=-1,
                 full_screen_resize=None, resize_resolution=None, ffmpeg_cc=False, verbose=False,
                 ignore_framerate_mismatch=False):
        self.fs_width = -1 if full_screen_resize is None else resize_resolution[0]
        self.fs_height = -1 if full_screen_resize is None else resize_resolution[1]
        if test_fname.endswith(".yuv") and reference_fname.endswith(".yuv"):
            self.reader = video_reader_yuv
        else:
            self.reader = video_reader if ffmpeg_cc else video_reader_yuv_pytorch
        self.reference_vidr = None
        self.test_vidr = None
        self.reference_fname, self.test_fname = reference_fname, test_fname
        self.in_frames = frames
        self.full_screen_resize, self.resize_resolution = full_screen_resize, resize_resolution
        self.ffmpeg_cc, self.verbose = ffmpeg_cc, verbose
        self.fps = fps
        self.ignore_framerate_mismatch = ignore_framerate_mismatch
        super().__init__(display_photometry=display_photometry, config_paths=config_paths)

    def get_frame_count(self):
        self.init_readers()
        return self.frames

    def init_readers(self):
        if self.reference_vidr is not None:
            return
        kw = dict(resize_fn=self.full_screen_resize, resize_width=self.fs_width, resize_height=self.fs_height,
                  verbose=self.verbose)
        self.reference_vidr = self.reader(self.reference_fname, self.in_frames, **kw)
        self.test_vidr = self.reader(self.test_fname, self.in_frames, **kw)
        tv, rv = self.test_vidr, self.reference_vidr
        if tv.frames == -1 and rv.frames == -1:
            logging.error("Neither test nor reference video contains meta-data with the number of frames. You need to "
                          "pass '--count-frames' or specify it with '--nframes' argument.")
            raise vq_exception("Unknown number of frames")
        if not self.ignore_framerate_mismatch:
            if tv.frames == -1:
                self.frames = rv.frames
            elif rv.frames == -1:
                self.frames = tv.frames
            else:
                self.frames = min(tv.frames, rv.frames)
                if tv.frames != rv.frames:
                    logging.warning(f"Test and reference videos contain different number of frames ({tv.frames} and "
                                    f"{rv.frames}). Comparing {self.frames} frames.")
        for vr, what, name in ((tv, "Test", self.test_fname), (rv, "Reference", self.reference_fname)):
            logging.debug(f"{what} video '{name}':")
            rs_str = "" if self.full_screen_resize is None else f"->[{self.resize_resolution[0]}x{self.resize_resolution[1]}]"
            if not self.ignore_framerate_mismatch:
                self.fps = vr.avg_fps if self.fps is None else self.fps
                logging.debug(f"  [{vr.src_width}x{vr.src_height}]{rs_str}, colorspace: {vr.color_space}, color transfer: "
                              f"{vr.color_transfer}, fps: {self.fps}, pixfmt: {vr.in_pix_fmt}, frames: {self.frames}")
        if not self.ignore_framerate_mismatch and tv.avg_fps != rv.avg_fps:
            raise vq_exception(f"Test and reference videos have different frame rates: test is {tv.avg_fps} fps, "
                               f"reference is {rv.avg_fps} fps. Pass `--temp-resample` to resample to a common frame rate.")
        if tv.color_transfer == "smpte2084" and self.dm_photometry.EOTF != "PQ":
            logging.warning(f"Video color transfer function ({tv.color_transfer}) inconsistent with EOTF of the display "
                            f"model ({self.dm_photometry.EOTF})")

    def get_video_size(self):
        self.init_readers()
        if getattr(self.test_vidr, "resize_fn", None) is not None:
            return (self.test_vidr.resize_height, self.test_vidr.resize_width, self.frames)
        return (self.test_vidr.height, self.test_vidr.width, self.frames)

    def get_frames_per_second(self):
        self.init_readers()
        return self.fps

    def get_test_frame(self, frame, device, colorspace="Y"):
        self.init_readers()
        return self._get_frame(self.test_vidr, frame, device, colorspace)

    def get_reference_frame(self, frame, device, colorspace="Y"):
        self.init_readers()
        return self._get_frame(self.reference_vidr, frame, device, colorspace)

    def yuv_readers(self):
        """(test reader, reference reader, first frame) when both deliver planar YUV of one format at the size the
        metric sees and nothing was read yet (the fused path reads the pipes itself, forward only); else None."""
        self.init_readers()
        tv, rv = self.test_vidr, self.reference_vidr
        if not (isinstance(tv, yuv_frame_decoder) and isinstance(rv, yuv_frame_decoder) and tv.same_format(rv)):
            return None
        for vr in (tv, rv):
            if vr.resize_fn is not None and (vr.resize_height != vr.height or vr.resize_width != vr.width):
                return None
            if vr.curr_frame != -1:
                return None
        return tv, rv, 0

    def _get_frame(self, vid_reader, frame, device, colorspace):
        self.init_readers()
        if frame != vid_reader.curr_frame + 1:
            raise vq_exception("Video can be currently only read frame-by-frame. Random access not implemented.")
        frame_np = vid_reader.get_frame()
        if frame_np is None:
            raise vq_exception(f'Could not read frame {frame} of "{vid_reader.fname}". Try passing "--count-frames" or '
                               '"-nframes".')
        return self._prepare_frame(frame_np, device, vid_reader.unpack, colorspace)

    def _prepare_frame(self, frame_np, device, unpack_fn, colorspace="Y"):
        frame_t = reshuffle_dims(unpack_fn(frame_np, device), in_dims="HWC", out_dims="BCFHW")
        return self.apply_dm_and_color_transform(frame_t, colorspace)


def safe_floor(x):
    """floor() that tolerates a value a rounding error below an integer (video_source_file.py:328-330)."""
    x_f = math.floor(x)
    return x_f if (x - x_f) < (1 - 1e-6) else x_f + 1


class video_source_temp_resample_file(video_source_video_file):
    """Test and reference with different (constant) frame rates: both are resampled in time, by frame repetition, to a
    common rate (`--temp-resample`, video_source_file.py:478-541).  Frames are pulled one by one (plugin path)."""

    max_fps = 166  # upsample to at most this rate

    def __init__(self, test_fname, reference_fname, display_photometry="sdr_4k_30", config_paths=[], frames=-1,
                 full_screen_resize=None, resize_resolution=None, ffmpeg_cc=False, verbose=False):
        super().__init__(test_fname, reference_fname, display_photometry=display_photometry, config_paths=config_paths,
                         frames=frames, full_screen_resize=full_screen_resize, resize_resolution=resize_resolution,
                         ffmpeg_cc=ffmpeg_cc, verbose=verbose, ignore_framerate_mismatch=True)
        super().init_readers()
        test_fps, ref_fps = self.test_vidr.avg_fps, self.reference_vidr.avg_fps
        cls = type(self)
        if test_fps > cls.max_fps or ref_fps > cls.max_fps:
            raise vq_exception(f"Maximum resample fps ({cls.max_fps}) is smaller than the fps of the test ({test_fps}) or "
                               f"reference video ({ref_fps}). Increase maximum resample fps, e.g, by passing "
                               f"`--temp-resample {max(test_fps, ref_fps)}`")
        if test_fps % 1 == 0 and ref_fps % 1 == 0:  # an integer common multiple, if it is not too high
            self.resample_fps = min(test_fps * ref_fps / math.gcd(int(test_fps), int(ref_fps)), cls.max_fps)
        else:
            self.resample_fps = cls.max_fps
        test_rs = int(self.test_vidr.frames * self.resample_fps / test_fps)
        ref_rs = int(self.reference_vidr.frames * self.resample_fps / ref_fps)
        if self.test_vidr.frames == -1:
            frames_rs = ref_rs
        elif self.reference_vidr.frames == -1:
            frames_rs = test_rs
        else:
            frames_rs = min(test_rs, ref_rs)
        self.frames = frames_rs if frames < 0 else frames
        logging.info(f"Test fps: {test_fps}; reference fps: {ref_fps}. Resampling videos to {self.resample_fps} frames per "
                     f"second. {self.frames} frames will be processed.")
        if test_rs != ref_rs:
            logging.warning(f"Test and reference videos contain different number of frames after resampling ({test_rs} and "
                            f"{ref_rs}). Comparing {self.frames} frames.")
        self.cache_ind = [-1, -1]
        self.cache_frame = [None, None]

    def get_frames_per_second(self):
        return self.resample_fps

    def yuv_readers(self):
        return None  # frames repeat: the clip is not a plain run of stream frames

    def _get_frame(self, vid_reader, frame, device, colorspace):
        frame_ind = int(safe_floor((frame + 0.5) * vid_reader.avg_fps / self.resample_fps))
        ce = 0 if vid_reader is self.test_vidr else 1
        if self.cache_ind[ce] != frame_ind:
            self.cache_ind[ce] = frame_ind
            self.cache_frame[ce] = super()._get_frame(vid_reader, frame_ind, device=device, colorspace=colorspace)
        return self.cache_frame[ce]
