"""Raw planar .yuv clips as a video source (pycvvdp/video_source_yuv.py), with the YUV->RGB conversion on
the GPU.

File-name metadata, frame geometry and the public classes follow the reference (`decode_video_props`,
`create_yuv_fname`, `YUVReader`, `video_source_yuv_file`).  The arithmetic of
`YUVReader._fixed2float_upscale` / `get_frame_rgb_tensor` -- limited-range unpack, bilinear chroma
upsampling, YCbCr->RGB, clip -- lives in the CUDA front end (csrc/cvvdp_kernels.cuh: yuv_fetch_rgb):
`cvvdp.predict_video_source` recognises a `video_source_yuv_file` and streams the memory-mapped frames
straight into the fused temporal kernel (1.5 bytes per pixel over PCIe for 8-bit 4:2:0 instead of 12
for fp32 RGB); `get_frame_rgb_tensor` / `get_test_frame` use the same device function for callers that
pull single frames.  The optional full-screen resize (torch.nn.functional.interpolate in the reference) is the
`k_resize` kernel, applied between the YCbCr matrix and the display model; resized sources go through the
frame-by-frame plugin path of the metric.
"""
import logging
import os
import re

import numpy as np
import torch

from . import _native as N
from .video_source import reshuffle_dims, video_source_dm

# YCbCr -> RGB coefficients, R = Y + c0 Cr, G = Y + c1 Cb + c2 Cr, B = Y + c3 Cb.  "2020" / "709": what the reference
# uses for raw .yuv files (video_source_yuv.py:167-177); "bt709": what it uses for frames decoded by ffmpeg
# (video_source_file.py:268-277; BT.2020nc streams share the "2020" row).
YCBCR2RGB = {"2020": (1.47460, -0.16455, -0.57135, 1.88140), "709": (1.402, -0.344136, -0.714136, 1.772),
             "bt709": (1.5748, -0.1873, -0.4681, 1.8556)}


def decode_video_props(fname):
    """Video properties encoded in a .yuv file name (video_source_yuv.py:7-63)."""
    vprops = {"width": 1920, "height": 1080, "fps": 24, "bit_depth": 8, "color_space": "709", "chroma_ss": "420"}
    bname = os.path.splitext(os.path.basename(fname))[0]
    res_match = re.compile(r"(\d+)x(\d+)p?(\d+)?")
    for field in bname.split("_"):
        if res_match.match(field):
            nums = re.findall(r"\d+", field)
            if len(nums) < 2 or len(nums) > 3:
                raise ValueError("Cannot decode the resolution")
            vprops["width"], vprops["height"] = int(nums[0]), int(nums[1])
            if len(nums) == 3:
                vprops["fps"] = int(nums[2])
        elif field.endswith("fps"):
            vprops["fps"] = float(field[:-3])
        elif field in ("444", "420", "422"):
            vprops["chroma_ss"] = field
        elif field in ("10", "10b", "10bit"):
            vprops["bit_depth"] = 10
        elif field in ("8", "8b", "8bit"):
            vprops["bit_depth"] = 8
        elif field in ("2020", "709"):
            vprops["color_space"] = field
        elif field in ("bt709", "sdr"):
            vprops["color_space"] = "709"
        elif field in ("ct2020", "pq2020", "hdr"):
            vprops["color_space"] = "2020"
    return vprops


def create_yuv_fname(basename, vprops):
    """video_source_yuv.py:66-74"""
    fps = vprops["fps"]
    fps = round(fps, 3) if round(fps) != fps else int(fps)
    return (f"{basename}_{vprops['width']}x{vprops['height']}_{vprops['bit_depth']}b_{vprops['chroma_ss']}_"
            f"{vprops['color_space']}_{fps}fps.yuv")


class yuv_frame_decoder:
    """Geometry of planar YUV frames and their decoding on the device (shared by the memory-mapped .yuv reader and
    the ffmpeg-pipe reader of video_source_file.py).  Subclasses set width, height, chroma_ss, bit_depth and
    `ycbcr_row` (a key of YCBCR2RGB) and call `_init_geometry`."""

    def _init_geometry(self):
        self.y_pixels = int(self.width * self.height)
        self.y_shape = (self.height, self.width)
        if self.chroma_ss == "444":
            self.frame_pixels = self.y_pixels * 3
            self.uv_shape = self.y_shape
        elif self.chroma_ss == "420":
            self.frame_pixels = self.y_pixels * 3 // 2
            self.uv_shape = (self.height // 2, self.width // 2)
        elif self.chroma_ss == "422":
            self.frame_pixels = self.y_pixels * 2
            self.uv_shape = (self.height, self.width // 2)
        else:
            raise RuntimeError(f"Unsupported chroma subsampling {self.chroma_ss}")
        self.uv_pixels = self.uv_shape[0] * self.uv_shape[1]
        self.dtype = np.uint16 if self.bit_depth > 8 else np.uint8
        self.frame_bytes = self.frame_pixels * (2 if self.bit_depth > 8 else 1)
        self._ctx = None

    # ---- native description --------------------------------------------------------------------
    def native_yuv(self) -> N.Yuv:
        y = N.Yuv()
        y.chroma, y.bit_depth = int(self.chroma_ss), int(self.bit_depth)
        for i, c in enumerate(YCBCR2RGB[self.ycbcr_row]):
            y.coef[i] = c
        return y

    def native_dtype(self):
        return N.DTYPE_U16 if self.bit_depth > 8 else N.DTYPE_U8

    def same_format(self, other):
        return all(getattr(self, k) == getattr(other, k) for k in ("width", "height", "chroma_ss", "bit_depth", "ycbcr_row"))

    def _device_ctx(self, device):
        """(context, torch device, stream) for single-frame decoding; the display is a pass-through."""
        from . import cvvdp_metric as cm
        from .display_model import vvdp_display_photo_eotf
        mock = cm._mock_library
        device = torch.device("cpu") if mock is not None else torch.device(device)
        if mock is None and (device.type != "cuda" or not torch.cuda.is_available()):
            raise RuntimeError("colorvideovdp_b200 needs a CUDA device (no CPU fallback)")
        key = 0 if mock is not None else (device.index if device.index is not None else torch.cuda.current_device())
        if self._ctx is None or self._ctx[0] != (key, mock):
            params, lut = cm._default_native_inputs()
            ctx = N.Context(params, lut, key, library=mock)
            ctx.set_display(vvdp_display_photo_eotf(1.0, contrast=1.0, EOTF="linear").native_display(passthrough=True))
            self._ctx = ((key, mock), ctx)
        stream = None if mock is not None else torch.cuda.current_stream(device).cuda_stream
        return self._ctx[1], device, stream

    def decode_frame(self, raw, device, resize=None):
        """One raw planar frame (1-D numpy array of frame_pixels samples) -> display-encoded RGB [H,W,3] in 0..1 on
        `device`: limited-range unpack, bilinear chroma upsampling, YCbCr matrix, clip (k_frontend), then the
        optional resize=(mode, height, width) (k_resize)."""
        ctx, device, stream = self._device_ctx(device)
        raw = np.array(raw, copy=True).reshape(-1)  # (file mappings and pipe buffers are read-only)
        if raw.size != self.frame_pixels:
            raise RuntimeError(f"a frame has {self.frame_pixels} samples, got {raw.size}")
        t = torch.from_numpy(raw.view(np.int16) if raw.dtype == np.uint16 else raw).to(device)
        clip = N.Clip()
        clip.data = t.data_ptr()
        clip.stride[0], clip.stride[2] = 0, self.frame_pixels
        clip.frame0, clip.n_frames = 0, 1
        out = torch.empty((3, self.height, self.width), dtype=torch.float32, device=device)
        ctx.frontend_yuv(clip, self.native_yuv(), 1, self.height, self.width, self.native_dtype(), 0, N.CS_RGB_LINEAR,
                         out.data_ptr(), stream)
        if resize is not None and (resize[1] != self.height or resize[2] != self.width):
            out = resize_planes(ctx, out, resize[1], resize[2], resize[0], stream)
        if stream is not None:
            torch.cuda.current_stream(device).synchronize()  # `t` is released on return
        return out.permute(1, 2, 0)


def resize_planes(ctx, planes, out_h, out_w, mode, stream):
    """[C,H,W] fp32 -> [C,out_h,out_w], clipped to 0..1 (F.interpolate(size=, mode=).clip(0, 1) of the reference)."""
    planes = planes.contiguous()
    Cc, H, W = planes.shape
    dst = torch.empty((Cc, int(out_h), int(out_w)), dtype=torch.float32, device=planes.device)
    ctx.resize(planes.data_ptr(), dst.data_ptr(), Cc, H, W, int(out_h), int(out_w), mode, True, stream)
    return dst


class YUVReader(yuv_frame_decoder):
    """Memory-mapped planar .yuv file (video_source_yuv.py:77-233)."""

    def __init__(self, file_name):
        self.file_name = file_name
        if not os.path.isfile(file_name):
            raise FileNotFoundError("File {} not found".format(file_name))
        vprops = decode_video_props(file_name)
        self.width, self.height = vprops["width"], vprops["height"]
        self.avg_fps = vprops["fps"]
        self.color_space = vprops["color_space"]
        self.chroma_ss = vprops["chroma_ss"]
        self.bit_depth = vprops["bit_depth"]
        self.ycbcr_row = "2020" if self.color_space == "2020" else "709"
        self._init_geometry()
        self.frames = int(os.stat(file_name).st_size / self.frame_bytes)
        self.mm = None

    def get_frame_count(self):
        return int(self.frames)

    def _map(self):
        if self.mm is None:
            self.mm = np.memmap(self.file_name, self.dtype, mode="r")
        return self.mm

    def get_frame_yuv(self, frame_index):
        if frame_index < 0 or frame_index >= self.frames:
            raise RuntimeError("The frame index is outside the range of available frames")
        mm = self._map()
        o = int(frame_index * self.frame_pixels)
        Y = mm[o:o + self.y_pixels]
        u = mm[o + self.y_pixels:o + self.y_pixels + self.uv_pixels]
        v = mm[o + self.y_pixels + self.uv_pixels:o + self.y_pixels + 2 * self.uv_pixels]
        return (np.reshape(Y, self.y_shape, "C"), np.reshape(u, self.uv_shape, "C"), np.reshape(v, self.uv_shape, "C"))

    def fileno(self):
        """A read-only descriptor of the file, for the library to pread() frame windows from (opened on first use)."""
        if getattr(self, "_fd", None) is None:
            self._fd = os.open(self.file_name, os.O_RDONLY)
        return self._fd

    def __del__(self):
        fd = getattr(self, "_fd", None)
        if fd is not None:
            try:
                os.close(fd)
            except OSError:
                pass
            self._fd = None

    def frames_window(self, first, count):
        """Frames [first, first+count) as one contiguous host array [count * frame_pixels] -- a view of the file
        mapping, no copy."""
        if first < 0 or first + count > self.frames:
            raise RuntimeError("The frame index is outside the range of available frames")
        return self._map()[first * self.frame_pixels:(first + count) * self.frame_pixels]

    def get_frame_rgb_tensor(self, frame_index, device):
        """Display-encoded RGB [H,W,3] in 0..1 on `device` (video_source_yuv.py:146-178)."""
        if frame_index < 0 or frame_index >= self.frames:
            raise RuntimeError("The frame index is outside the range of available frames")
        return self.decode_frame(self.frames_window(frame_index, 1), device)

    def __enter__(self):
        return self

    def __exit__(self, type, value, tb):
        self.mm = None


class video_reader_yuv(YUVReader):
    """A .yuv reader with the interface of the ffmpeg readers (video_source_yuv.py:233-261): sequential
    `get_frame` / `unpack`, optional resize."""

    def __init__(self, vidfile, frames=-1, resize_fn=None, resize_height=-1, resize_width=-1, verbose=False):
        super().__init__(vidfile)
        self.fname = vidfile
        self.src_width, self.src_height = self.width, self.height
        self.in_pix_fmt = "yuv" + self.chroma_ss + "p"
        self.resize_fn, self.resize_width, self.resize_height = resize_fn, resize_width, resize_height
        self.color_transfer = None
        if frames != -1:
            self.frames = min(self.frames, frames)
        self.curr_frame = -1

    def get_frame(self):
        self.curr_frame += 1
        return self.curr_frame

    def unpack(self, frame_index, device):
        if frame_index < 0 or frame_index >= self.frames:
            raise RuntimeError("The frame index is outside the range of available frames")
        rs = None if self.resize_fn is None else (self.resize_fn, self.resize_height, self.resize_width)
        return self.decode_frame(self.frames_window(frame_index, 1), device, resize=rs)

    def close(self):
        self.mm = None


class video_source_yuv_file(video_source_dm):
    """Test/reference pair of raw .yuv files (video_source_yuv.py:264-352)."""

    def __init__(self, test_fname, reference_fname, display_photometry="standard_4k", frames=-1, full_screen_resize=None,
                 resize_resolution=None, retain_aspect_ratio=False, verbose=False):
        self.reference_vidr = YUVReader(reference_fname)
        self.test_vidr = YUVReader(test_fname)
        self.total_frames = self.test_vidr.frames
        self.frames = self.total_frames if frames == -1 else min(self.total_frames, frames)
        self.offset = 0
        self.full_screen_resize = full_screen_resize
        if full_screen_resize is not None and full_screen_resize not in N.RESIZE_MODES:
            raise RuntimeError(f"unknown full_screen_resize '{full_screen_resize}'")
        if retain_aspect_ratio:  # video_source_yuv.py:275-282
            h, w = self.test_vidr.height, self.test_vidr.width
            if h / resize_resolution[1] * resize_resolution[0] <= w:
                resize_resolution = (resize_resolution[0], int(resize_resolution[0] / w * h))
            else:
                resize_resolution = (int(resize_resolution[1] / h * w), resize_resolution[1])
        self.resize_resolution = resize_resolution  # (width, height)
        super().__init__(display_photometry=display_photometry)
        for vr, name in ((self.test_vidr, test_fname), (self.reference_vidr, reference_fname)):
            rs_str = "" if full_screen_resize is None else f"->[{resize_resolution[0]}x{resize_resolution[1]}]"
            logging.debug(f"Video '{name}': [{vr.width}x{vr.height}]{rs_str}, colorspace: {vr.color_space}, "
                          f"EOTF: {self.dm_photometry.EOTF}, fps: {vr.avg_fps}, frames: {self.frames}")

    def get_video_size(self):
        if self.full_screen_resize is not None:
            return [self.resize_resolution[1], self.resize_resolution[0], self.frames]
        return [self.test_vidr.height, self.test_vidr.width, self.frames]

    def get_frames_per_second(self):
        return self.test_vidr.avg_fps

    def get_test_frame(self, frame, device, colorspace="Y"):
        return self._get_frame(self.test_vidr, frame, device, colorspace)

    def get_reference_frame(self, frame, device, colorspace="Y"):
        return self._get_frame(self.reference_vidr, frame, device, colorspace)

    def resizes(self, vid_reader):
        """The frames of `vid_reader` change size on the way in (video_source_yuv.py:333)."""
        return self.full_screen_resize is not None and (vid_reader.height != self.resize_resolution[1] or
                                                        vid_reader.width != self.resize_resolution[0])

    def yuv_readers(self):
        """(test reader, reference reader, first frame) when the raw frames can go straight into the fused temporal
        kernel -- same format, no resize, enough frames -- else None."""
        tr, rr = self.test_vidr, self.reference_vidr
        if not tr.same_format(rr) or self.resizes(tr) or self.resizes(rr):
            return None
        if self.frames > min(tr.frames, rr.frames) - self.offset:
            return None
        return tr, rr, self.offset

    def _get_frame(self, vid_reader, frame, device, colorspace="Y"):
        rs = None
        if self.resizes(vid_reader):
            rs = (self.full_screen_resize, self.resize_resolution[1], self.resize_resolution[0])
        if self.offset + frame < 0 or self.offset + frame >= vid_reader.frames:
            raise RuntimeError("The frame index is outside the range of available frames")
        RGB = vid_reader.decode_frame(vid_reader.frames_window(self.offset + frame, 1), device, resize=rs)
        RGB_bcfhw = reshuffle_dims(RGB, in_dims="HWC", out_dims="BCFHW")
        return self.apply_dm_and_color_transform(RGB_bcfhw, colorspace)

    def set_offset(self, offset: int):
        self.offset = offset

    def get_total_frames(self):
        return self.total_frames

    def set_num_frames(self, num_frames: int):
        if self.offset + num_frames > self.total_frames:
            logging.error(f"Cannot set num_frames={num_frames} because offset={self.offset} and "
                          f"total_frames={self.total_frames}. Clipping num_frames to {self.total_frames - self.offset}")
            num_frames = self.total_frames - self.offset
        self.frames = num_frames
