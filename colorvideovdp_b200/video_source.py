"""Frame sources: the `video_source` plugin surface of the reference (pycvvdp/video_source.py).

``video_source_array`` keeps the test/reference clips as torch tensors (a view in BCFHW order, no
copy).  The metric reads those tensors directly and fuses the display model into its CUDA front end;
``get_test_frame`` / ``get_reference_frame`` remain for callers that pull frames one by one.
"""
import logging

import numpy as np
import torch

from .display_model import vvdp_display_photometry


class video_source:
    """Abstract frame source (pycvvdp/video_source.py:17-78)."""

    def get_video_size(self):
        """(height, width, frames)"""
        raise NotImplementedError

    def get_frames_per_second(self) -> float:
        raise NotImplementedError

    def get_test_frame(self, frame, device, colorspace):
        raise NotImplementedError

    def get_reference_frame(self, frame, device, colorspace):
        raise NotImplementedError

    def check_if_valid(self, frame, target_colorspace):
        """NaN / Inf / photometric-scale warnings of pycvvdp/video_source.py:48-72."""
        if not hasattr(self, "warning_shown"):
            self.warning_shown = False
        if not self.warning_shown and bool(torch.isnan(frame).any()):
            self.warning_shown = True
            logging.warning("Image contains one or more NaN values")
        if not self.warning_shown and bool(torch.isinf(frame).any()):
            self.warning_shown = True
            logging.warning("Image contains one or more Inf values")
        if not hasattr(self, "first_frame"):
            self.first_frame = True
        if self.first_frame and not target_colorspace.startswith("display_encoded") and target_colorspace != "RGB2020pq":
            self.first_frame = False
            f_mean = float(frame[:, 0].mean())
            logging.debug(f"Content mean={f_mean}, max={float(frame[:, 0].max())}, min={float(frame[:, 0].min())}")
            if not self.warning_shown and f_mean <= 1:
                logging.warning("The mean color value is less than 1 - the image may not be scaled in absolute "
                                "photometric units!")

    def get_frame_count(self):
        return self.get_video_size()[2]

    def get_batch_size(self):
        return 1


def reshuffle_dims(T, in_dims: str, out_dims: str):
    """Permute / add singleton dimensions, e.g. "HWC" -> "BCFHW" (pycvvdp/video_source.py:120-162).
    Returns a view whenever torch can express the result as one."""
    in_dims, out_dims = in_dims.upper(), out_dims.upper()
    assert len(in_dims) == T.dim(), "The in_dims string must have as many characters as there are dimensions in T"
    kept = [d for d in out_dims if d in in_dims]
    for k in sorted((k for k, d in enumerate(in_dims) if d not in out_dims), reverse=True):
        assert T.shape[k] == 1, "Only the dimensions of size 1 can be skipped in the output"
        T = T.squeeze(dim=k)
    in_kept = [d for d in in_dims if d in out_dims]
    T_p = T.permute([in_kept.index(d) for d in kept])
    shape = [T_p.shape[kept.index(d)] if d in kept else 1 for d in out_dims]
    return T_p.reshape(shape)


def _as_tensor(video):
    """numpy / torch -> torch tensor without copying; uint16 is carried as an int16 bit pattern
    (pycvvdp/video_source.py:257-271)."""
    if isinstance(video, np.ndarray):
        if video.dtype == np.uint16:
            video = video.view(np.int16)
        return torch.from_numpy(video) if video.flags.writeable else torch.from_numpy(video.copy())
    if hasattr(torch, "uint16") and video.dtype == torch.uint16:
        return video.view(torch.int16)
    return video


class video_source_dm(video_source):
    """A source that applies a photometric display model (pycvvdp/video_source.py:204-222)."""

    def __init__(self, display_photometry="sdr_4k_30", config_paths=[]):
        if isinstance(display_photometry, str):
            self.dm_photometry = vvdp_display_photometry.load(display_photometry, config_paths)
        elif isinstance(display_photometry, vvdp_display_photometry):
            self.dm_photometry = display_photometry
        else:
            raise RuntimeError("display_model must be a string or fvvdp_display_photometry subclass")

    def apply_dm_and_color_transform(self, frame, target_colorspace):
        I = self.dm_photometry.source_2_target_colorspace(frame, target_colorspace)
        self.check_if_valid(I, target_colorspace)
        return I


class video_source_array(video_source_dm):
    """Test/reference clips held in numpy arrays or torch tensors (pycvvdp/video_source.py:234-346)."""

    def __init__(self, test_video, reference_video, fps, dim_order="BCFHW", display_photometry="sdr_4k_30",
                 config_paths=[]):
        super().__init__(display_photometry=display_photometry, config_paths=config_paths)
        if test_video.shape != reference_video.shape:
            ind = dim_order.find("B")
            rest_t = [s for k, s in enumerate(test_video.shape) if k != ind]
            rest_r = [s for k, s in enumerate(reference_video.shape) if k != ind]
            # only the batch dimension may differ, and only as a singleton (video_source.py:247-252)
            if not (ind >= 0 and len(test_video.shape) == len(reference_video.shape) and rest_t == rest_r
                    and (test_video.shape[ind] == 1 or reference_video.shape[ind] == 1)):
                raise RuntimeError("Test and reference image/video tensors must be exactly the same shape")
        if len(dim_order) != len(test_video.shape):
            raise RuntimeError('Input tensor much have exactly as many dimensions as there are characters in the '
                               '"dims" parameter')
        test_video = reshuffle_dims(_as_tensor(test_video), in_dims=dim_order, out_dims="BCFHW")
        reference_video = reshuffle_dims(_as_tensor(reference_video), in_dims=dim_order, out_dims="BCFHW")
        B, Cc, F, H, W = test_video.shape
        if fps == 0 and F > 1:
            raise RuntimeError("When passing video sequences, you must set frames_per_second parameter")
        if Cc not in (1, 3):
            raise RuntimeError("The content must have either 1 or 3 color channels.")
        for v in (test_video, reference_video):
            if v.dtype not in (torch.float32, torch.float16, torch.int16, torch.uint8):
                raise RuntimeError(f"Only uint8, uint16 and float32 is currently supported. {v.dtype} encountered.")
        self.fps = fps
        self.is_video = fps > 0
        self.is_color = Cc == 3
        self.test_video = test_video
        self.reference_video = reference_video

    def get_frames_per_second(self):
        return self.fps

    def get_video_size(self):
        sh = self.test_video.shape
        return (sh[3], sh[4], sh[2])

    def get_batch_size(self):
        return max(self.test_video.shape[0], self.reference_video.shape[0])

    def get_test_frame(self, frame, device, colorspace):
        return self._get_frame(self.test_video, frame, device, colorspace)

    def get_reference_frame(self, frame, device, colorspace):
        return self._get_frame(self.reference_video, frame, device, colorspace)

    def _get_frame(self, from_array, frame, device, colorspace):
        """One [B,C,1,H,W] fp32 frame in `colorspace` on `device` (pycvvdp/video_source.py:320-346)."""
        sl = from_array[:, :, frame:frame + 1].to(device)
        if sl.dtype == torch.int16:
            sl = (sl.to(torch.int32) & 0xFFFF).to(torch.float32) / 65535
        elif sl.dtype == torch.uint8:
            sl = sl.to(torch.float32) / 255
        elif sl.dtype == torch.float16:
            sl = sl.to(torch.float32)
        return self.apply_dm_and_color_transform(sl, colorspace)
