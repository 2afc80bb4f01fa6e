"""ctypes binding of libcvvdp_b200.so (the C ABI in include/cvvdp_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU fallback:
a missing library or a machine without a CUDA device raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcvvdp_b200.so")

MAX_BANDS = 16
MAX_FILTER_LEN = 129
CSF_LUT_N = 32

DTYPE_U8, DTYPE_U16, DTYPE_F16, DTYPE_F32 = 0, 1, 2, 3
EOTF_SRGB, EOTF_PQ, EOTF_LINEAR, EOTF_HLG, EOTF_GAMMA, EOTF_NONE = 0, 1, 2, 3, 4, 5
PAD_REPLICATE, PAD_SYMMETRIC = 0, 1
HEATMAP_NONE, HEATMAP_RAW, HEATMAP_THRESHOLD, HEATMAP_SUPRATHRESHOLD = 0, 1, 2, 3
CS_DKLD65, CS_RGB_LINEAR, CS_XYZ, CS_LMS2006 = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [("mask_p", C.c_float), ("mask_c", C.c_float), ("mask_q", C.c_float * 4),
                ("xcm_weights", C.c_float * 16), ("beta", C.c_float), ("beta_t", C.c_float),
                ("beta_tch", C.c_float), ("beta_sch", C.c_float), ("sensitivity_correction", C.c_float),
                ("jod_a", C.c_float), ("jod_exp", C.c_float), ("image_int", C.c_float),
                ("ch_chrom_w", C.c_float), ("ch_trans_w", C.c_float), ("baseband_weight", C.c_float * 4),
                ("d_max", C.c_float), ("sigma_tf", C.c_float * 4), ("beta_tf", C.c_float * 4),
                ("pu_dilate", C.c_float)]


class CsfLut(C.Structure):
    _fields_ = [("L_bkg", C.c_float * CSF_LUT_N), ("rho", C.c_float * CSF_LUT_N),
                ("logS", ((C.c_float * CSF_LUT_N) * CSF_LUT_N) * 4)]


class Display(C.Structure):
    _fields_ = [("eotf", C.c_int32), ("gamma", C.c_float), ("Y_peak", C.c_float), ("contrast", C.c_float),
                ("E_ambient", C.c_float), ("k_refl", C.c_float), ("exposure", C.c_float),
                ("rgb2xyz", C.c_float * 9), ("ppd", C.c_float)]


class Yuv(C.Structure):
    _fields_ = [("chroma", C.c_int32), ("bit_depth", C.c_int32), ("coef", C.c_float * 4)]


class Job(C.Structure):
    _fields_ = [("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("n_frames", C.c_int32),
                ("fps", C.c_float), ("in_channels", C.c_int32), ("dtype", C.c_int32), ("padding", C.c_int32),
                ("heatmap", C.c_int32), ("max_block_frames", C.c_int32), ("workspace_limit_bytes", C.c_int64),
                ("yuv", Yuv), ("features", C.c_int32), ("prefiltered", C.c_int32)]


class PlanInfo(C.Structure):
    _fields_ = [("n_bands", C.c_int32), ("n_channels", C.c_int32), ("filter_len", C.c_int32),
                ("block_frames", C.c_int32), ("rho_band", C.c_float * MAX_BANDS),
                ("band_height", C.c_int32 * MAX_BANDS), ("band_width", C.c_int32 * MAX_BANDS),
                ("filters", (C.c_float * MAX_FILTER_LEN) * 4), ("workspace_bytes", C.c_int64)]


class Clip(C.Structure):
    _fields_ = [("data", C.c_void_p), ("stride", C.c_int64 * 5), ("frame0", C.c_int32), ("n_frames", C.c_int32)]


# every symbol include/cvvdp_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cvvdp_b200_abi_version": (C.c_int, []),
    "cvvdp_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "cvvdp_b200_create": (C.c_int, [C.POINTER(Params), C.POINTER(CsfLut), C.c_int, C.POINTER(C.c_void_p)]),
    "cvvdp_b200_destroy": (None, [C.c_void_p]),
    "cvvdp_b200_set_display": (C.c_int, [C.c_void_p, C.POINTER(Display)]),
    "cvvdp_b200_plan": (C.c_int, [C.c_void_p, C.POINTER(Job), C.POINTER(PlanInfo)]),
    "cvvdp_b200_process_device": (C.c_int, [C.c_void_p, C.POINTER(Clip), C.POINTER(Clip), C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvvdp_b200_process_host": (C.c_int, [C.c_void_p, C.POINTER(Clip), C.POINTER(Clip), C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]),
    "cvvdp_b200_process_files": (C.c_int, [C.c_void_p, C.POINTER(Clip), C.POINTER(Clip), C.c_int, C.c_int, C.c_int64,
                                           C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cvvdp_b200_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvvdp_b200_pool_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p]),
    "cvvdp_b200_frontend": (C.c_int, [C.c_void_p, C.POINTER(Clip), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvvdp_b200_frontend_yuv": (C.c_int, [C.c_void_p, C.POINTER(Clip), C.POINTER(Yuv), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cvvdp_b200_resize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p]),
    "cvvdp_b200_launch_count": (C.c_int64, [C.c_void_p]),
    "cvvdp_b200_band_strip_width": (C.c_int, [C.c_void_p, C.c_int]),
    "cvvdp_b200_input_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "cvvdp_b200_temporal_filters": (C.c_int, [C.c_void_p, C.c_float, C.POINTER(C.c_float)]),
    "cvvdp_b200_feature_layout": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "cvvdp_b200_set_feature_output": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cvvdp_b200_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "cvvdp_b200_profile_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
}

KERNEL_KINDS = ["temporal", "reduce", "band", "baseband", "finalize", "heatmap", "pool", "frontend", "features"]


class KernelStat(C.Structure):
    _fields_ = [("kind", C.c_int32), ("level", C.c_int32), ("launches", C.c_int32), ("total_ms", C.c_float),
                ("algo_bytes", C.c_double)]

ABI_VERSION = 6
RESIZE_MODES = {"nearest": 0, "bilinear": 1, "bicubic": 2, "area": 3}  # run_cvvdp.py:100


class InputReport(C.Structure):
    _fields_ = [("out_of_range", C.c_int64), ("nan", C.c_int64), ("inf", C.c_int64), ("first_frame_sum", C.c_double)]


class NativeError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen the native library and bind every ABI symbol.  Raises when it is missing."""
    if not os.path.isfile(path):
        raise NativeError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). colorvideovdp_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.cvvdp_b200_abi_version() != ABI_VERSION:
        raise NativeError("ABI version mismatch between the Python binding and the native library")
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


class Context:
    """Owns one cvvdp_b200_ctx."""

    def __init__(self, params: Params, lut: CsfLut, device_index: int, library=None):
        self._lib = library if library is not None else lib()
        self._h = C.c_void_p()
        rc = self._lib.cvvdp_b200_create(C.byref(params), C.byref(lut), int(device_index), C.byref(self._h))
        if rc != 0:
            msg = self._lib.cvvdp_b200_last_error(None)
            self._h = C.c_void_p()
            raise NativeError(f"cvvdp_b200_create failed ({rc}): {msg.decode() if msg else ''}")

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.cvvdp_b200_last_error(self._h)
            raise NativeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cvvdp_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_display(self, disp: Display):
        self._check(self._lib.cvvdp_b200_set_display(self._h, C.byref(disp)), "set_display")

    def plan(self, job: Job) -> PlanInfo:
        info = PlanInfo()
        self._check(self._lib.cvvdp_b200_plan(self._h, C.byref(job), C.byref(info)), "plan")
        return info

    def process_device(self, test: Clip, ref: Clip, f0, f1, q_ptr, hm_ptr, stream):
        self._check(self._lib.cvvdp_b200_process_device(self._h, C.byref(test), C.byref(ref), int(f0), int(f1),
                                                        q_ptr, hm_ptr, stream), "process_device")

    def process_host(self, test: Clip, ref: Clip, f0, f1, q_ptr, hm_ptr):
        self._check(self._lib.cvvdp_b200_process_host(self._h, C.byref(test), C.byref(ref), int(f0), int(f1),
                                                      q_ptr, hm_ptr), "process_host")

    def process_files(self, test: Clip, ref: Clip, fd_test, fd_ref, off_test, off_ref, f0, f1, q_ptr, hm_ptr):
        self._check(self._lib.cvvdp_b200_process_files(self._h, C.byref(test), C.byref(ref), fd_test, fd_ref, off_test,
                                                       off_ref, f0, f1, q_ptr, hm_ptr), "process_files")

    def pool(self, q_ptr, B, Cc, F, L, jod_ptr):
        self._check(self._lib.cvvdp_b200_pool(self._h, q_ptr, B, Cc, F, L, jod_ptr), "pool")

    def pool_device(self, q_ptr, B, Cc, F, L, jod_ptr, stream):
        self._check(self._lib.cvvdp_b200_pool_device(self._h, q_ptr, B, Cc, F, L, jod_ptr, stream), "pool_device")

    def frontend(self, src: Clip, B, cin, H, W, dtype, frame, colorspace, dst_ptr, flags_ptr, stream):
        self._check(self._lib.cvvdp_b200_frontend(self._h, C.byref(src), B, cin, H, W, dtype, frame, colorspace,
                                                  dst_ptr, flags_ptr, stream), "frontend")

    def frontend_yuv(self, src: Clip, yuv: Yuv, B, H, W, dtype, frame, colorspace, dst_ptr, stream):
        self._check(self._lib.cvvdp_b200_frontend_yuv(self._h, C.byref(src), C.byref(yuv), B, H, W, dtype, frame,
                                                      colorspace, dst_ptr, stream), "frontend_yuv")

    def resize(self, src_ptr, dst_ptr, channels, H, W, out_h, out_w, mode, clip01, stream):
        """`mode`: a key of RESIZE_MODES (the reference's --full-screen-resize choices)."""
        if mode not in RESIZE_MODES:
            raise NativeError(f"resize: unknown interpolation '{mode}'")
        self._check(self._lib.cvvdp_b200_resize(self._h, src_ptr, dst_ptr, channels, H, W, out_h, out_w,
                                                RESIZE_MODES[mode], 1 if clip01 else 0, stream), "resize")

    def launch_count(self):
        return int(self._lib.cvvdp_b200_launch_count(self._h))

    def temporal_filters(self, fps):
        """[4][n] taps (A-sust, RG, YV, A-trans) at `fps`, as a list of lists."""
        buf = (C.c_float * (4 * MAX_FILTER_LEN))()
        n = self._lib.cvvdp_b200_temporal_filters(self._h, float(fps), buf)
        if n < 0:
            raise NativeError(f"temporal_filters: error {n}")
        return [[buf[c * MAX_FILTER_LEN + k] for k in range(n)] for c in range(4)]

    def feature_layout(self, band):
        """(ph, pw, feature_size, float offset) of the feature tensor of `band` (band == n_bands: total floats)."""
        ph, pw, fs, off = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self._lib.cvvdp_b200_feature_layout(self._h, int(band), C.byref(ph), C.byref(pw), C.byref(fs),
                                                        C.byref(off)), "feature_layout")
        return ph.value, pw.value, fs.value, off.value

    def set_feature_output(self, ptr):
        self._check(self._lib.cvvdp_b200_set_feature_output(self._h, ptr), "set_feature_output")

    def band_strip_width(self, level):
        return int(self._lib.cvvdp_b200_band_strip_width(self._h, int(level)))

    def input_stats(self, reset=True):
        """Validation counters of the fused front end since the last reset (synchronises the device)."""
        rep = InputReport()
        self._check(self._lib.cvvdp_b200_input_stats(self._h, C.byref(rep), 1 if reset else 0), "input_stats")
        return rep

    def profile_enable(self, on=True):
        self._check(self._lib.cvvdp_b200_profile_enable(self._h, 1 if on else 0), "profile_enable")

    def profile_read(self):
        """[{kind, level, launches, total_ms, algo_bytes}] since the last read (synchronises the device)."""
        buf = (KernelStat * 128)()
        n = C.c_int(0)
        self._check(self._lib.cvvdp_b200_profile_read(self._h, buf, 128, C.byref(n)), "profile_read")
        return [dict(kind=KERNEL_KINDS[buf[i].kind], level=buf[i].level, launches=buf[i].launches,
                     total_ms=float(buf[i].total_ms), algo_bytes=float(buf[i].algo_bytes)) for i in range(n.value)]
