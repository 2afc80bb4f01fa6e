"""ColorVideoVDP metric object -- drop-in for ``pycvvdp.cvvdp`` on the B200 hot path.

Mirrors the public surface of the reference class (pycvvdp/cvvdp_metric.py:108-1218): constructor
arguments, ``predict`` / ``predict_video_source``, ``set_display_model``, ``do_pooling_and_jods``,
``get_info_string``, ``write_features_to_json`` and the ``stats`` dictionary.  All arithmetic of the
hot loop runs in the native CUDA library (csrc/) behind the C ABI of include/cvvdp_b200.h; this file
is host-side plumbing only and raises if the library or a CUDA device is missing.
"""
import ctypes
import json
import logging
import math

import numpy as np
import torch

from . import _native as N
from . import utils
from .display_model import vvdp_display_geometry, vvdp_display_photo_eotf, vvdp_display_photometry
from .video_source import video_source_array
from .vq_metric import register_metric, vq_exception, vq_metric

_TORCH_DTYPES = {torch.uint8: N.DTYPE_U8, torch.int16: N.DTYPE_U16, torch.float16: N.DTYPE_F16,
                 torch.float32: N.DTYPE_F32}

# Test hook: tests/ may inject a mock native library (tests/emu) to exercise the host logic and the
# kernel logic on machines without a GPU.  Never set by the package itself.
_mock_library = None


def _set_mock_library_for_tests(library):
    global _mock_library
    _mock_library = library


def _native_inputs(config_paths):
    """cvvdp_parameters.json + csf_lut_*.json -> the parameter structs of the C ABI
    (replaces cvvdp.load_config, cvvdp_metric.py:146-229, and castleCSF.__init__, csf.py:8-25)."""
    parameters_file = utils.config_files.find("cvvdp_parameters.json", config_paths)
    par = utils.json2dict(parameters_file)
    unsupported = []
    if par.get("masking_model") != "mult-mutual":
        unsupported.append(f"masking_model={par.get('masking_model')}")
    if par.get("contrast") != "weber_g1":
        unsupported.append(f"contrast={par.get('contrast')}")
    if par.get("dclamp_type") != "soft":
        unsupported.append(f"dclamp_type={par.get('dclamp_type')}")
    if par.get("xchannel_masking") != "on":
        unsupported.append("xchannel_masking=off")
    if par.get("temp_filter", "default") != "default" or "block_channels" in par:
        unsupported.append("temp_filter/block_channels")
    if unsupported:
        raise NotImplementedError("colorvideovdp_b200 implements the shipped ColorVideoVDP model only; "
                                  "unsupported parameters: " + ", ".join(unsupported))
    p = N.Params()
    for key in ("mask_p", "mask_c", "beta", "beta_t", "beta_tch", "beta_sch", "sensitivity_correction", "jod_a",
                "jod_exp", "image_int", "ch_chrom_w", "ch_trans_w", "d_max", "pu_dilate"):
        setattr(p, key, float(par[key]))
    for key, n in (("mask_q", 4), ("xcm_weights", 16), ("baseband_weight", 4), ("sigma_tf", 4), ("beta_tf", 4)):
        vals = list(par[key])
        if len(vals) != n:
            raise RuntimeError(f"parameter {key} must have {n} entries")
        arr = getattr(p, key)
        for i in range(n):
            arr[i] = float(vals[i])
    lut_file = utils.config_files.find(f"csf_lut_{par['csf']}.json", config_paths)
    lut_json = utils.json2dict(lut_file)
    lut = N.CsfLut()
    n = N.CSF_LUT_N
    if len(lut_json["L_bkg"]) != n or len(lut_json["rho"]) != n:
        raise RuntimeError("CSF LUT must be 32x32")
    for i in range(n):
        lut.L_bkg[i] = float(lut_json["L_bkg"][i])
        lut.rho[i] = float(lut_json["rho"][i])
    om = lut_json["omega"]
    names = [f"o{om[0]}_c1", f"o{om[0]}_c2", f"o{om[0]}_c3", f"o{om[1]}_c1"]  # csf.py:17-23
    for c, name in enumerate(names):
        table = np.asarray(lut_json[name], dtype=np.float32)
        ctypes.memmove(lut.logS[c], np.ascontiguousarray(table).ctypes.data, n * n * 4)
    return p, lut, par, parameters_file


_default_cache = None


def _default_native_inputs():
    global _default_cache
    if _default_cache is None:
        _default_cache = _native_inputs([])[:2]
    return _default_cache


_band_freq_cache = {}


def _band_frequencies(width, height, ppd):
    """rho_band as the reference reports it in stats (lpyr_dec.py:18-52, cvvdp_metric.py:685-686)."""
    key = (int(width), int(height), float(ppd))
    if key not in _band_freq_cache:
        if len(_band_freq_cache) > 64:
            _band_freq_cache.clear()
        _band_freq_cache[key] = _band_frequencies_uncached(width, height, ppd)
    return _band_freq_cache[key].copy()


def _band_frequencies_uncached(width, height, ppd):
    max_levels = int(np.floor(np.log2(min(height, width)))) - 1
    bands = np.concatenate([[1.0], np.power(2.0, -np.arange(0.0, 14.0)) * 0.3228], 0) * ppd / 2.0
    invalid = np.nonzero(bands <= 0.2)[0]
    max_band = max_levels if invalid.size == 0 else invalid[0]
    n = int(np.clip(max_band + 1, 0, max_levels))
    freqs = np.array([1.0] + [0.3228 * 2.0 ** (-f) for f in range(n)]) * ppd / 2.0
    freqs[n] = 0.1
    return freqs


def _as_float32(t):
    """video_source_array._get_frame's unpacking (video_source.py:324-342) for a whole clip."""
    if t.dtype == torch.uint8:
        return t.to(torch.float32) / 255
    if t.dtype == torch.int16:
        return (t.to(torch.int32) & 0xFFFF).to(torch.float32) / 65535
    return t.to(torch.float32)


def _clip_of(t, batch, frame0=0):
    c = N.Clip()
    c.data = t.data_ptr()
    st = t.stride()
    for i in range(5):
        c.stride[i] = st[i]
    if t.shape[0] == 1 and batch > 1:
        c.stride[0] = 0  # singleton batch broadcast (video_source.py:247-252)
    c.frame0, c.n_frames = frame0, t.shape[2]
    return c


class cvvdp(vq_metric):
    def __init__(self, display_name="standard_4k", display_photometry=None, display_geometry=None, config_paths=[],
                 heatmap=None, quiet=False, device=None, temp_padding="replicate", use_checkpoints=False,
                 dump_channels=None, gpu_mem=None):
        self.quiet = quiet
        self.heatmap = heatmap
        self.temp_padding = temp_padding
        self.gpu_mem = gpu_mem  # GB of workspace the engine may use
        self.training_mode = False
        assert heatmap in ["threshold", "supra-threshold", "raw", "none", None], "Unknown heatmap type"
        self.do_heatmap = (self.heatmap is not None) and (self.heatmap != "none")
        if use_checkpoints:
            raise NotImplementedError("use_checkpoints (autograd through the metric) is outside the scope of the "
                                      "CUDA hot path")
        if dump_channels:
            raise NotImplementedError("dump_channels (debug video dumps) is outside the scope of the CUDA hot path")
        self.dump_channels = None
        if _mock_library is not None:
            self.device = torch.device("cpu")  # tests only: mock device shares the host address space
        else:
            if device is None:
                device = torch.device("cuda")
            device = torch.device(device)
            if device.type != "cuda" or not torch.cuda.is_available():
                raise RuntimeError("colorvideovdp_b200 runs on a CUDA device only: there is no CPU fallback "
                                   f"(requested device '{device}', torch.cuda.is_available()="
                                   f"{torch.cuda.is_available()})")
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            self.device = device
        self._plan_key = None
        self._info = None
        self.set_display_model(display_name, display_photometry=display_photometry,
                               display_geometry=display_geometry, config_paths=config_paths)
        self.load_config(config_paths)

    # ------------------------------------------------------------------------------------------
    def train(self, do_training=True):
        self.training_mode = do_training

    def load_config(self, config_paths):
        params, lut, par, self.parameters_file = _native_inputs(config_paths)
        logging.debug(f"Loading ColorVideoVDP parameters from '{self.parameters_file}'")
        self._params = params
        self.version = par["version"]
        # the calibrated values, exposed under the reference's attribute names
        dev = self.device
        for key in ("mask_p", "mask_c", "beta", "beta_t", "beta_tch", "beta_sch", "sensitivity_correction", "jod_a",
                    "jod_exp", "image_int", "ch_chrom_w", "ch_trans_w", "d_max", "mask_q", "xcm_weights",
                    "baseband_weight", "sigma_tf", "beta_tf"):
            setattr(self, key, torch.as_tensor(par[key], dtype=torch.float32, device=dev))
        self.pu_dilate = par["pu_dilate"]
        self.masking_model = par["masking_model"]
        self.contrast = par["contrast"]
        dev_index = 0 if self.device.type != "cuda" else self.device.index
        self._ctx = N.Context(params, lut, dev_index, library=_mock_library)
        self._plan_key = None

    def set_display_model(self, display_name="standard_4k", display_photometry=None, display_geometry=None,
                          config_paths=[]):
        if display_photometry is None:
            self.display_photometry = vvdp_display_photometry.load(display_name, config_paths)
            self.display_name = display_name
        else:
            self.display_photometry = display_photometry
            self.display_name = getattr(display_photometry, "short_name", "unspecified")
        if display_geometry is None:
            self.display_geometry = vvdp_display_geometry.load(display_name, config_paths)
        else:
            self.display_geometry = display_geometry
        self.pix_per_deg = self.display_geometry.get_ppd()
        self._plan_key = None

    def predict(self, test_cont, reference_cont, dim_order="BCFHW", frames_per_second=0):
        test_vs = video_source_array(test_cont, reference_cont, frames_per_second, dim_order=dim_order,
                                     display_photometry=self.display_photometry)
        return self.predict_video_source(test_vs)

    def loss(self, test_cont, reference_cont, dim_order="BCFHW", frames_per_second=0):
        """10 - JOD (cvvdp_metric.py:294-298), as a value.  The reference differentiates through its torch ops; this
        engine has forward kernels only, so content that requires a gradient is refused instead of silently returning a
        constant."""
        for c in (test_cont, reference_cont):
            if isinstance(c, torch.Tensor) and c.requires_grad:
                raise NotImplementedError("colorvideovdp_b200 has no backward kernels: loss() cannot be differentiated "
                                          "(detach the tensors to get the value, or use the reference for optimisation)")
        Q_jod, _ = self.predict(test_cont, reference_cont, dim_order=dim_order, frames_per_second=frames_per_second)
        return 10.0 - Q_jod

    # ------------------------------------------------------------------------------------------
    def _plan(self, B, H, W, F, fps, cin, dtype_id, photo, yuv=None, features=False, prefiltered=False):
        """(Re)build the native plan when the job or the display changed.  photo=None: frames already
        are DKLd65 (plugin sources), the front end passes them through."""
        if self.temp_padding not in ("replicate", "symmetric"):
            raise RuntimeError(f'Unknown padding method "{self.temp_padding}"')
        if photo is None:
            disp = vvdp_display_photo_eotf(1.0, contrast=1.0, EOTF="linear").native_display(self.pix_per_deg,
                                                                                           passthrough=True)
        else:
            disp = photo.native_display(self.pix_per_deg)
        hm = {"raw": N.HEATMAP_RAW, "threshold": N.HEATMAP_THRESHOLD,
              "supra-threshold": N.HEATMAP_SUPRATHRESHOLD}[self.heatmap] if self.do_heatmap else N.HEATMAP_NONE
        key = (B, H, W, F, float(fps), cin, dtype_id, self.temp_padding, hm, bytes(disp), self.gpu_mem,
               bytes(yuv) if yuv is not None else None, bool(features), bool(prefiltered))
        if key != self._plan_key:
            self._ctx.set_display(disp)
            job = N.Job(batch=B, height=H, width=W, n_frames=F, fps=float(fps), in_channels=cin, dtype=dtype_id,
                        padding=N.PAD_REPLICATE if self.temp_padding == "replicate" else N.PAD_SYMMETRIC,
                        heatmap=hm, max_block_frames=0,
                        workspace_limit_bytes=int(self.gpu_mem * 1e9) if self.gpu_mem else 0)
            if yuv is not None:
                job.yuv = yuv
            job.features = 1 if features else 0
            job.prefiltered = 1 if prefiltered else 0
            self._info = self._ctx.plan(job)
            self._plan_key = key
        return self._info

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def _needed_frames(self, f0, f1, fl, F):
        """Clip frames the temporal stage reads for outputs [f0, f1) (cvvdp_metric.py:501-529)."""
        lo, hi = f1 - 1, f1
        for t in range(f0 - (fl - 1), f1):
            s = t
            if s < 0:
                s = 0 if self.temp_padding == "replicate" else self._get_symmetric_frame_index(s, F)
            lo, hi = min(lo, s), max(hi, s + 1)
        return lo, hi

    def _get_symmetric_frame_index(self, frame_ind, frame_count):
        """cvvdp_metric.py:445-450"""
        is_even = (math.floor((abs(frame_ind) - 1) / (frame_count - 1)) % 2) == 0
        if is_even:
            return ((abs(frame_ind) - 1) % (frame_count - 1)) + 1
        return frame_ind % (frame_count - 1)

    def compute_q_per_ch(self, vid_source, frame_range=None):
        """The hot loop: Q_per_ch [B,C,F,L] (torch, on the metric's device; frames outside `frame_range`
        are zero) and the raw heat map (fp16 [1,1,F,H,W] on the device, or None)."""
        H, W, F = vid_source.get_video_size()
        B = vid_source.get_batch_size()
        if B > 1 and self.do_heatmap:
            raise vq_exception("Heatmaps not supported when batches are used")
        fps = vid_source.get_frames_per_second() if F > 1 else 0
        f0, f1 = (0, F) if frame_range is None else frame_range
        if F > 1 and getattr(vid_source, "is_temporally_filtered", False):
            return self._run_prefiltered(vid_source, B, H, W, F, fps, f0, f1)
        fast = type(vid_source) is video_source_array and type(vid_source.dm_photometry) is vvdp_display_photo_eotf
        if fast:
            return self._run_arrays(vid_source, B, H, W, F, fps, f0, f1)
        from .video_source_yuv import video_source_yuv_file
        from .video_source_file import video_source_video_file
        if type(vid_source) in (video_source_yuv_file, video_source_video_file) and \
                type(vid_source.dm_photometry) is vvdp_display_photo_eotf:
            readers = vid_source.yuv_readers()
            if readers is not None:
                return self._run_yuv(vid_source, readers, H, W, F, fps, f0, f1)
        return self._run_plugin(vid_source, B, H, W, F, fps, f0, f1)

    yuv_chunk_bytes = 1 << 30  # host window (test + reference) handed to one process_host call of the YUV path

    def _run_yuv(self, vs, readers, H, W, F, fps, f0, f1):
        """Raw planar YUV frames (a memory-mapped .yuv file or an ffmpeg pipe) go straight into the fused temporal
        kernel, which unpacks, upsamples the chroma and applies the YCbCr matrix on the fly
        (video_source_yuv.py:146-233, video_source_file.py:261-324).  The clip is walked in windows of whole plan
        blocks, in increasing frame order, so that a pipe is read once and host memory stays bounded."""
        tr, rr, off = readers
        info = self._plan(1, H, W, F, fps, 3, tr.native_dtype(), vs.dm_photometry, tr.native_yuv())
        fl = info.filter_len
        nb = max(1, info.block_frames)
        per = max(1, int(self.yuv_chunk_bytes // (2 * tr.frame_bytes * nb))) * nb
        per = max(per, -(-fl // nb) * nb)
        pin = self.device.type == "cuda"
        Qh = torch.zeros((1, info.n_channels, F, info.n_bands), dtype=torch.float32, pin_memory=pin)
        hmh = None
        if self.do_heatmap:
            make = torch.empty if (f0 == 0 and f1 == F) else torch.zeros
            hmh = make((1, self._hm_channels(), F, H, W), dtype=torch.float16, pin_memory=pin)
        cur = f0
        while cur < f1:
            end = min(cur + per, f1)
            wlo, whi = self._needed_frames(cur, end, fl, F)
            from_files = all(hasattr(r, "fileno") for r in (tr, rr))
            # .yuv files: the library reads the window from the descriptors; pipes: from the readers' buffers
            windows = [] if from_files else [r.frames_window(off + wlo, whi - wlo) for r in (tr, rr)]
            clips = []
            for k in range(2):
                c = N.Clip()
                c.data = None if from_files else windows[k].ctypes.data
                c.stride[0], c.stride[2] = 0, tr.frame_pixels
                c.frame0, c.n_frames = wlo, whi - wlo
                clips.append(c)
            # process_host / process_files return the whole [1,C,F,L] array, zero outside [cur, end): collect the windows
            Qw = Qh if (cur == f0 and end == f1) else torch.empty_like(Qh)
            hm_ptr = hmh.data_ptr() if hmh is not None else None
            if from_files:
                for r in (tr, rr):
                    if off + whi > r.frames:
                        raise RuntimeError("The frame index is outside the range of available frames")
                self._ctx.process_files(clips[0], clips[1], tr.fileno(), rr.fileno(), (off + wlo) * tr.frame_bytes,
                                        (off + wlo) * rr.frame_bytes, cur, end, Qw.data_ptr(), hm_ptr)
            else:
                self._ctx.process_host(clips[0], clips[1], cur, end, Qw.data_ptr(), hm_ptr)
            if Qw is not Qh:
                Qh[:, :, cur:end] = Qw[:, :, cur:end]
            cur = end
        return Qh, hmh

    def report_input_problems(self, vid_source, n_pixels, saw_frame0):
        """The input checks the reference makes frame by frame on the host, from the counters the fused front end
        keeps on the device: 'Pixel outside the valid range 0-1' (display_model.py:335-337), NaN and photometric-scale
        warnings (video_source.py:48-72, once per video source) and the failure on non-finite differences
        (cvvdp_metric.py:906-907: a NaN pixel always reaches D).  Synchronises the device; call after the hot loop."""
        rep = self._ctx.input_stats(reset=True)
        shown = getattr(vid_source, "warning_shown", False) if vid_source is not None else False
        if rep.out_of_range:
            logging.warning("Pixel outside the valid range 0-1")
        if rep.nan and not shown:
            shown = True
            logging.warning("Image contains one or more NaN values")
        if saw_frame0 and n_pixels > 0 and not shown:
            f_mean = rep.first_frame_sum / n_pixels
            logging.debug(f"Content mean={f_mean}")
            if f_mean <= 1:
                logging.warning("The mean color value is less than 1 - the image may not be scaled in absolute "
                                "photometric units!")
        if vid_source is not None:
            vid_source.warning_shown = shown
            vid_source.first_frame = False
        if rep.nan:
            raise AssertionError("Must not be nan")

    def _alloc_outputs(self, B, C, F, L, H, W, whole=False):
        Q = torch.zeros((B, C, F, L), dtype=torch.float32, device=self.device)
        hm = None
        if self.do_heatmap:  # every frame of the evaluated range is written by the kernels; zero only what is not
            make = torch.empty if whole else torch.zeros
            hm = make((1, self._hm_channels(), F, H, W), dtype=torch.float16, device=self.device)
        return Q, hm

    def _hm_channels(self):
        return 1 if self.heatmap == "raw" else 3  # cvvdp_metric.py:343

    def _run_arrays(self, vs, B, H, W, F, fps, f0, f1):
        """Fast path: raw clip tensors, display model fused into the CUDA front end."""
        if vs.dm_photometry is not self.display_photometry and vs.dm_photometry != self.display_photometry:
            logging.warning("video source and metric use different display models; using the video source's")
        return self.q_per_ch_from_tensors(vs.test_video, vs.reference_video, F, fps, (f0, f1), 0, vs.dm_photometry)

    def q_per_ch_from_tensors(self, test, ref, n_frames_total, frames_per_second, frame_range=None, first_frame=0,
                              photometry=None, _resident=None):
        """Q_per_ch for BCFHW tensors that hold clip frames [first_frame, first_frame + test.shape[2]) of a
        clip with `n_frames_total` frames (a window is enough as long as it covers the temporal support
        of `frame_range`).  Device tensors go through process_device, host tensors through the streaming
        process_host path.  Returns (Q_per_ch [B,C,F_total,L] zero outside frame_range, raw heat map)."""
        photo = photometry if photometry is not None else self.display_photometry
        if not isinstance(photo, vvdp_display_photo_eotf):
            raise RuntimeError("the fused front end needs a vvdp_display_photo_eotf display model")
        if test.dtype != ref.dtype:  # the reference unpacks frame by frame and accepts a mix (video_source.py:324-342)
            test, ref = _as_float32(test), _as_float32(ref)
        B = max(test.shape[0], ref.shape[0])
        H, W, F = test.shape[3], test.shape[4], int(n_frames_total)
        if B > 1 and self.do_heatmap:
            raise vq_exception("Heatmaps not supported when batches are used")
        fps = frames_per_second if F > 1 else 0
        f0, f1 = (0, F) if frame_range is None else frame_range
        info = self._plan(B, H, W, F, fps, test.shape[1], _TORCH_DTYPES[test.dtype], photo)
        C, L = info.n_channels, info.n_bands
        resident = (test.device.type != "cpu" or ref.device.type != "cpu") if _resident is None else _resident
        if resident:  # (_resident is a test hook: on the mock device every tensor is a CPU tensor)
            test, ref = test.to(self.device), ref.to(self.device)
            Q, hm = self._alloc_outputs(B, C, F, L, H, W, whole=(f0 == 0 and f1 == F))
            self._ctx.process_device(_clip_of(test, B, first_frame), _clip_of(ref, B, first_frame), f0, f1,
                                     Q.data_ptr(), hm.data_ptr() if hm is not None else None, self._stream())
            return Q, hm
        # host clips: streamed upload overlapped with compute inside the native library
        pin = self.device.type == "cuda"
        Qh = torch.zeros((B, C, F, L), dtype=torch.float32, pin_memory=pin)
        hmh = None
        if self.do_heatmap:
            make = torch.empty if (f0 == 0 and f1 == F) else torch.zeros
            hmh = make((1, self._hm_channels(), F, H, W), dtype=torch.float16, pin_memory=pin)
        try:
            self._ctx.process_host(_clip_of(test, B, first_frame), _clip_of(ref, B, first_frame), f0, f1,
                                   Qh.data_ptr(), hmh.data_ptr() if hmh is not None else None)
        except N.NativeError as e:
            if "densely" not in str(e):
                raise
            test, ref = test.contiguous(), ref.contiguous()  # exotic layout: normalise to BCFHW first
            self._ctx.process_host(_clip_of(test, B, first_frame), _clip_of(ref, B, first_frame), f0, f1,
                                   Qh.data_ptr(), hmh.data_ptr() if hmh is not None else None)
        return Qh, hmh

    def get_temporal_filters(self, frames_per_s):
        """cvvdp_metric.py:1057-1092: (list of the four filters [N] on the metric's device, omega_bands [0, 5]).
        The taps come from the native planner, i.e. they are the ones the temporal kernel applies."""
        F = [torch.tensor(row, dtype=torch.float32, device=self.device) for row in self._ctx.temporal_filters(frames_per_s)]
        return F, torch.as_tensor([0.0, 5.0], device=self.device)

    def extract_features(self, vid_source):
        """Per-band feature tensors for the ML heads -- cvvdp_ml_base.extract_features with
        cvvdp_feature_pooling (pycvvdp/cvvdp_ml_metric.py:78-106, 206-298, 302-352): a list with one
        tensor [B, F, ph, pw, C, 6] per band on the metric's device, last axis (mean_T, var_T, mean_R,
        var_R, mean_D, var_D) of |T_f| S, |R_f| S and D over ceil(ppd) x ceil(ppd) patches, and None for
        the heat map (the reference's ML metrics do not produce one).  The regression heads themselves
        need downloaded weights and are not part of this package."""
        if self.do_heatmap:
            raise vq_exception("Currently cvvdp-ml metrics do not produce heatmaps")
        if not (type(vid_source) is video_source_array and type(vid_source.dm_photometry) is vvdp_display_photo_eotf):
            raise NotImplementedError("extract_features needs a video_source_array with a vvdp_display_photo_eotf display")
        H, W, F = vid_source.get_video_size()
        B = vid_source.get_batch_size()
        fps = vid_source.get_frames_per_second() if F > 1 else 0
        test, ref = vid_source.test_video, vid_source.reference_video
        if test.dtype != ref.dtype:
            raise RuntimeError("Test and reference must have the same dtype")
        info = self._plan(B, H, W, F, fps, test.shape[1], _TORCH_DTYPES[test.dtype], vid_source.dm_photometry, features=True)
        C, L = info.n_channels, info.n_bands
        layout = [self._ctx.feature_layout(bb) for bb in range(L + 1)]
        buf = torch.zeros((layout[L][3],), dtype=torch.float32, device=self.device)
        test, ref = test.to(self.device), ref.to(self.device)
        Q, _ = self._alloc_outputs(B, C, F, L, H, W)
        self._ctx.set_feature_output(buf.data_ptr())
        try:
            self._ctx.process_device(_clip_of(test, B, 0), _clip_of(ref, B, 0), 0, F, Q.data_ptr(), None, self._stream())
        finally:
            self._ctx.set_feature_output(None)
        feats = []
        for bb in range(L):
            ph, pw, _, off = layout[bb]
            feats.append(buf[off:layout[bb + 1][3]].view(B, F, ph, pw, C, 6))
        return feats, None

    def _run_prefiltered(self, vs, B, H, W, F, fps, f0, f1):
        """Sources that deliver the four temporal channels themselves (cvvdp_metric.py:470-488, e.g. an
        eye-motion model): frames come as 'DKLd65_trans' [B,4,1,H,W], reference before test, and enter the CUDA
        path as level 0 -- the temporal filter is bypassed."""
        vs._frames_pulled_through_plugin = True
        info = self._plan(B, H, W, F, fps, 4, N.DTYPE_F32, None, prefiltered=True)
        Q, hm = self._alloc_outputs(B, info.n_channels, F, info.n_bands, H, W, whole=(f0 == 0 and f1 == F))
        nb = info.block_frames
        for cur in range(f0, f1, nb):
            end = min(cur + nb, f1)
            fr_r, fr_t = [], []
            for f in range(cur, end):
                fr_r.append(vs.get_reference_frame(f, device=self.device, colorspace="DKLd65_trans").to(
                    device=self.device, dtype=torch.float32))
                fr_t.append(vs.get_test_frame(f, device=self.device, colorspace="DKLd65_trans").to(
                    device=self.device, dtype=torch.float32))
            win_t, win_r = torch.cat(fr_t, dim=2), torch.cat(fr_r, dim=2)
            if win_t.shape[1] != 4 or win_r.shape[1] != 4:
                raise RuntimeError("a temporally filtered source must deliver four channels ('DKLd65_trans')")
            self._ctx.process_device(_clip_of(win_t, B, cur), _clip_of(win_r, B, cur), cur, end, Q.data_ptr(),
                                     hm.data_ptr() if hm is not None else None, self._stream())
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()  # the windows are released below
        return Q, hm

    def _run_plugin(self, vs, B, H, W, F, fps, f0, f1):
        """Any third-party video_source: frames are pulled through get_*_frame(..., 'DKLd65'), strictly in
        increasing order and reference before test, and enter the CUDA path at the temporal stage (the
        display model is the plugin's business)."""
        fl = 1 if F == 1 else int(math.ceil(0.250 * fps / 2) * 2) + 1  # cvvdp_metric.py:1059
        cache_t, cache_r = {}, {}
        vs._frames_pulled_through_plugin = True  # the source's own get_*_frame did the input checks

        def fetch(f):
            if f not in cache_r:
                cache_r[f] = vs.get_reference_frame(f, device=self.device, colorspace="DKLd65").to(
                    device=self.device, dtype=torch.float32)
                cache_t[f] = vs.get_test_frame(f, device=self.device, colorspace="DKLd65").to(
                    device=self.device, dtype=torch.float32)
            return cache_t[f], cache_r[f]

        first_lo, _ = self._needed_frames(f0, f0 + 1, fl, F)
        cin = fetch(first_lo)[1].shape[1]
        info = self._plan(B, H, W, F, fps, cin, N.DTYPE_F32, None)
        assert info.filter_len == fl
        Q, hm = self._alloc_outputs(B, info.n_channels, F, info.n_bands, H, W)
        nb = info.block_frames
        cur = f0
        while cur < f1:
            end = min(cur + nb, f1)
            lo, hi = self._needed_frames(cur, end, fl, F)
            frames = [fetch(f) for f in range(lo, hi)]
            win_t = torch.cat([fr[0] for fr in frames], dim=2)
            win_r = torch.cat([fr[1] for fr in frames], dim=2)
            self._ctx.process_device(_clip_of(win_t, B, lo), _clip_of(win_r, B, lo), cur, end, Q.data_ptr(),
                                     hm.data_ptr() if hm is not None else None, self._stream())
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()  # the windows are released below
            cur = end
            if cur < f1:
                keep_lo, _ = self._needed_frames(cur, min(cur + nb, f1), fl, F)
                for f in [k for k in cache_r if k < keep_lo]:
                    cache_t.pop(f), cache_r.pop(f)
        return Q, hm

    # ------------------------------------------------------------------------------------------
    def predict_video_source(self, vid_source, frame_range=None):
        """JOD and statistics for the clip of `vid_source` (cvvdp_metric.py:304-441).  `frame_range`
        restricts the frames evaluated by this process (frame sharding, see distributed.py); the JOD then
        only covers those frames unless the caller all-reduces stats['Q_per_ch'] first."""
        H, W, F = vid_source.get_video_size()
        Q_per_ch, heatmap = self.compute_q_per_ch(vid_source, frame_range)
        if isinstance(getattr(vid_source, "dm_photometry", None), vvdp_display_photo_eotf) and \
                not hasattr(vid_source, "_frames_pulled_through_plugin"):
            self.report_input_problems(vid_source, vid_source.get_batch_size() * H * W,
                                       frame_range is None or frame_range[0] == 0)
        Q_jod = self.do_pooling_and_jods(Q_per_ch if frame_range is None else Q_per_ch[:, :, frame_range[0]:frame_range[1]])
        stats = self._make_stats(Q_per_ch, heatmap, vid_source, H, W, F)
        return (Q_jod.squeeze(), stats)

    def _make_stats(self, Q_per_ch, heatmap, vid_source, H, W, F):
        stats = {}
        stats["Q_per_ch"] = Q_per_ch.detach().cpu().numpy()
        stats["rho_band"] = _band_frequencies(W, H, self.pix_per_deg)
        stats["frames_per_second"] = vid_source.get_frames_per_second()
        stats["width"] = W
        stats["height"] = H
        stats["N_frames"] = F
        if self.do_heatmap:  # raw [1,1,F,H,W] or coloured [1,3,F,H,W], fp16, from the native kernels, on the CPU
            hm = heatmap.detach()
            if hm.device.type == "cuda":  # through pinned memory: a pageable .cpu() of a 4K clip's map runs at a few GB/s
                host = torch.empty(hm.shape, dtype=hm.dtype, pin_memory=True)
                host.copy_(hm, non_blocking=True)
                torch.cuda.current_stream(hm.device).synchronize()
                hm = host
            stats["heatmap"] = hm
        return stats

    def get_ch_weights(self, no_channels):
        w = torch.stack([torch.as_tensor(1.0, device=self.ch_chrom_w.device), self.ch_chrom_w, self.ch_chrom_w,
                         self.ch_trans_w])
        return w[0:no_channels].view(1, -1, 1, 1)

    def do_pooling_and_jods(self, Q_per_ch):
        """Pool Q_per_ch [B,C,F,L] over bands, channels and frames and map to JOD
        (cvvdp_metric.py:610-643); runs in the native pooling kernel."""
        if isinstance(Q_per_ch, np.ndarray):
            Q_per_ch = torch.from_numpy(np.ascontiguousarray(Q_per_ch, dtype=np.float32))
        Q = Q_per_ch.detach().to(torch.float32).contiguous()
        B, C, F, L = Q.shape
        if Q.device == self.device and self.device.type == "cuda":
            jod = torch.empty((B,), dtype=torch.float32, device=self.device)
            self._ctx.pool_device(Q.data_ptr(), B, C, F, L, jod.data_ptr(), self._stream())
            return jod.squeeze()
        Qh = Q.cpu()
        jod = torch.empty((B,), dtype=torch.float32)
        self._ctx.pool(Qh.data_ptr(), B, C, F, L, jod.data_ptr())
        return jod.to(self.device).squeeze()

    def met2jod(self, Q):
        """cvvdp_metric.py:646-658 (tiny, element-wise; host tensors or device tensors alike)."""
        Q = torch.as_tensor(Q)
        Q_t = 0.1
        jod_a, jod_exp = float(self.jod_a), float(self.jod_exp)
        jod_a_p = jod_a * (Q_t ** (jod_exp - 1.0))
        return torch.where(Q <= Q_t, 10.0 - jod_a_p * Q, 10.0 - jod_a * Q.clamp(min=Q_t) ** jod_exp)

    def full_name(self):
        return "ColorVideoVDP"

    def quality_unit(self):
        return "JOD"

    def short_name(self):
        return "cvvdp"

    def get_info_string(self):
        if self.display_name.startswith("standard_"):
            standard_str = self.display_name
        else:
            standard_str = f"custom-display: {self.display_name}"
        L_black, L_refl = self.display_photometry.get_black_level()
        return f'"{self.full_name()} v{self.version}, {self.pix_per_deg:.4g} [pix/deg], ' \
               f'Lpeak={self.display_photometry.get_peak_luminance():.5g}, ' \
               f'Lblack={L_black:.4g}, Lrefl={L_refl:.4g} [cd/m^2], ({standard_str})"'

    def write_features_to_json(self, stats, dest_fname):
        """cvvdp_metric.py:1112-1127: per-channel/per-band features `t{c}_b{b}` for calibration/."""
        Q_per_ch = stats["Q_per_ch"]
        fmap = {}
        for key, value in stats.items():
            if key not in ["Q_per_ch", "heatmap"]:
                fmap[key] = value.tolist() if isinstance(value, np.ndarray) else value
        for cc in range(Q_per_ch.shape[1]):
            for bb in range(Q_per_ch.shape[3]):
                fmap[f"t{cc}_b{bb}"] = Q_per_ch[:, cc, :, bb].tolist()
        with open(dest_fname, "w", encoding="utf-8") as f:
            json.dump(fmap, f, ensure_ascii=False, indent=4)

    def distogram_array(self, stats):
        """Numeric part of export_distogram (cvvdp_metric.py:1160-1170): per channel/frame/band JOD loss."""
        Q = torch.as_tensor(stats["Q_per_ch"], dtype=torch.float32).clone()
        if Q.shape[0] != 1:
            raise vq_exception("Exporting distograms in batch mode is not supported")
        ch_no = Q.shape[1]
        Q[:, :, :, -1] *= self.baseband_weight[0:ch_no].cpu().view(-1, 1)
        Q *= self.get_ch_weights(ch_no).cpu() * ch_no
        return (10.0 - self.met2jod(Q)).numpy()

    def export_distogram(self, stats, fname, jod_max=None, base_size=6):
        try:
            import matplotlib.pyplot as plt  # noqa: F401
        except Exception:
            raise RuntimeError("matplotlib is missing. Please install it before exporting distograms.")
        raise NotImplementedError("plotting is outside the scope of the CUDA hot path; use distogram_array(stats)")


register_metric(cvvdp)
