"""Raw planar .yuv clips as a video source (pycvvdp/video_source_yuv.py), with the YUV->RGB conversion on
the GPU.

File-name metadata, frame geometry and the public classes follow the reference (`decode_video_props`,
`create_yuv_fname`, `YUVReader`, `video_source_yuv_file`).  The arithmetic of
`YUVReader._fixed2float_upscale` / `get_frame_rgb_tensor` -- limited-range unpack, bilinear chroma
upsampling, YCbCr->RGB, clip -- lives in the CUDA front end (csrc/cvvdp_kernels.cuh: yuv_fetch_rgb):
`cvvdp.predict_video_source` recognises a `video_source_yuv_file` and streams the memory-mapped frames
straight into the fused temporal kernel (1.5 bytes per pixel over PCIe for 8-bit 4:2:0 instead of 12
for fp32 RGB); `get_frame_rgb_tensor` / `get_test_frame` use the same device function for callers that
pull single frames.  Full-screen resizing is not implemented (out of scope of this path).
"""
import logging
import os
import re

import numpy as np
import torch

from . import _native as N
from .video_source import reshuffle_dims, video_source_dm

# YCbCr -> RGB coefficients used by the reference for raw .yuv files (video_source_yuv.py:167-177):
# R = Y + c0 Cr, G = Y + c1 Cb + c2 Cr, B = Y + c3 Cb
YCBCR2RGB = {"2020": (1.47460, -0.16455, -0.57135, 1.88140), "709": (1.402, -0.344136, -0.714136, 1.772)}


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


class YUVReader:
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
        self.frames = int(os.stat(file_name).st_size / self.frame_bytes)
        self.mm = None
        self._ctx = None

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

    # ---- native description --------------------------------------------------------------------
    def native_yuv(self) -> N.Yuv:
        y = N.Yuv()
        y.chroma, y.bit_depth = int(self.chroma_ss), int(self.bit_depth)
        for i, c in enumerate(YCBCR2RGB["2020" if self.color_space == "2020" else "709"]):
            y.coef[i] = c
        return y

    def native_dtype(self):
        return N.DTYPE_U16 if self.bit_depth > 8 else N.DTYPE_U8

    def frames_tensor(self, first, count):
        """Frames [first, first+count) as a [count, frame_pixels] CPU tensor (uint8 / int16 bit pattern)."""
        mm = self._map()
        a = np.array(mm[first * self.frame_pixels:(first + count) * self.frame_pixels], copy=True).reshape(count, self.frame_pixels)
        if a.dtype == np.uint16:
            a = a.view(np.int16)
        return torch.from_numpy(a)

    def get_frame_rgb_tensor(self, frame_index, device):
        """Display-encoded RGB [H,W,3] in 0..1 on `device` (video_source_yuv.py:146-178), via k_frontend."""
        if frame_index < 0 or frame_index >= self.frames:
            raise RuntimeError("The frame index is outside the range of available frames")
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
        ctx = self._ctx[1]
        raw = self.frames_tensor(frame_index, 1).to(device)
        clip = N.Clip()
        clip.data = raw.data_ptr()
        clip.stride[0], clip.stride[2] = 0, self.frame_pixels
        clip.frame0, clip.n_frames = 0, 1
        out = torch.empty((3, self.height, self.width), dtype=torch.float32, device=device)
        stream = None if mock is not None else torch.cuda.current_stream(device).cuda_stream
        ctx.frontend_yuv(clip, self.native_yuv(), 1, self.height, self.width, self.native_dtype(), 0, N.CS_RGB_LINEAR,
                         out.data_ptr(), stream)
        if stream is not None:
            torch.cuda.current_stream(device).synchronize()  # `raw` is released on return
        return out.permute(1, 2, 0)

    def __enter__(self):
        return self

    def __exit__(self, type, value, tb):
        self.mm = None


class video_source_yuv_file(video_source_dm):
    """Test/reference pair of raw .yuv files (video_source_yuv.py:264-352)."""

    def __init__(self, test_fname, reference_fname, display_photometry="standard_4k", frames=-1, full_screen_resize=None,
                 resize_resolution=None, retain_aspect_ratio=False, verbose=False):
        if full_screen_resize is not None:
            raise NotImplementedError("full_screen_resize is outside the scope of the CUDA hot path")
        self.reference_vidr = YUVReader(reference_fname)
        self.test_vidr = YUVReader(test_fname)
        self.total_frames = self.test_vidr.frames
        self.frames = self.total_frames if frames == -1 else min(self.total_frames, frames)
        self.offset = 0
        self.full_screen_resize = None
        self.resize_resolution = resize_resolution
        super().__init__(display_photometry=display_photometry)
        for vr, name in ((self.test_vidr, test_fname), (self.reference_vidr, reference_fname)):
            logging.debug(f"Video '{name}': [{vr.width}x{vr.height}], colorspace: {vr.color_space}, "
                          f"EOTF: {self.dm_photometry.EOTF}, fps: {vr.avg_fps}, frames: {self.frames}")

    def get_video_size(self):
        return [self.test_vidr.height, self.test_vidr.width, self.frames]

    def get_frames_per_second(self):
        return self.test_vidr.avg_fps

    def get_test_frame(self, frame, device, colorspace="Y"):
        return self._get_frame(self.test_vidr, frame, device, colorspace)

    def get_reference_frame(self, frame, device, colorspace="Y"):
        return self._get_frame(self.reference_vidr, frame, device, colorspace)

    def _get_frame(self, vid_reader, frame, device, colorspace="Y"):
        RGB = vid_reader.get_frame_rgb_tensor(self.offset + frame, device)
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
