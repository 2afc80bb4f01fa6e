"""Display models: the `display_model` plugin surface of the reference (pycvvdp/display_model.py).

Host-side objects only describe the display (JSON loading, black level, pixels per degree); the
arithmetic -- EOTF and RGB->DKL -- runs in the CUDA front end (csrc/cvvdp_kernels.cuh), either fused
into the temporal kernel (fast path of ``cvvdp.predict``) or through ``forward`` /
``source_2_target_colorspace`` below for callers that use the display model on its own.
"""
import logging
import math

import torch

from . import _native as N
from . import utils

_EOTF_IDS = {"sRGB": N.EOTF_SRGB, "PQ": N.EOTF_PQ, "linear": N.EOTF_LINEAR, "HLG": N.EOTF_HLG}
_COLORSPACE_IDS = {"DKLd65": N.CS_DKLD65, "XYZ": N.CS_XYZ, "LMS2006": N.CS_LMS2006}


class vvdp_display_photometry:
    """Base class (pycvvdp/display_model.py:110-276).  Subclasses implement ``forward``."""

    def __init__(self, source_colorspace="sRGB", config_paths=[]):
        colorspaces_file = utils.config_files.find("color_spaces.json", config_paths)
        colorspaces = utils.json2dict(colorspaces_file)
        if source_colorspace not in colorspaces:
            raise RuntimeError(f'Color space: "{source_colorspace}" not found in "{colorspaces_file}"')
        cs = colorspaces[source_colorspace]
        if "RGB2X" in cs:
            self.rgb2xyz_list = [cs["RGB2X"], cs["RGB2Y"], cs["RGB2Z"]]
        self.EOTF = cs["EOTF"]

    def forward(self, V):
        raise NotImplementedError

    def print(self):
        pass

    @classmethod
    def list_displays(cls, config_paths):
        models_file = utils.config_files.find("display_models.json", config_paths)
        logging.info(f"JSON file with display models: {models_file}")
        for display_name in utils.json2dict(models_file):
            vvdp_display_photometry.load(display_name, config_paths).print()

    @classmethod
    def load(cls, display_name, config_paths):
        models_file = utils.config_files.find("display_models.json", config_paths)
        models = utils.json2dict(models_file)
        if display_name not in models:
            logging.error(f"Display model: '{display_name}' not found in '{models_file}'")
            raise RuntimeError("Display model not found")
        model = models[display_name]
        Y_peak = model["max_luminance"]
        if "min_luminance" in model:
            contrast = Y_peak / model["min_luminance"]
        else:
            contrast = model.get("contrast", 500)
        obj = vvdp_display_photo_eotf(Y_peak, contrast=contrast, source_colorspace=model.get("colorspace", "sRGB"),
                                      E_ambient=model.get("E_ambient", 0), k_refl=model.get("k_refl", 0.005),
                                      name=display_name, exposure=model.get("exposure", 1),
                                      config_paths=config_paths)
        obj.full_name = model["name"]
        obj.short_name = display_name
        return obj

    def source_2_target_colorspace(self, I_src, target_colorspace):
        """pycvvdp/display_model.py:206-237 for the colour spaces the metric uses."""
        if isinstance(self, vvdp_display_photo_eotf) and target_colorspace in _COLORSPACE_IDS:
            return self._run_frontend(I_src, _COLORSPACE_IDS[target_colorspace])
        if target_colorspace.startswith("display_encoded") or target_colorspace in ("Y", "RGB709", "RGB2020",
                                                                                    "RGB2020pq", "logLMS_DKLd65"):
            raise NotImplementedError(f"target colour space '{target_colorspace}' is not used by ColorVideoVDP and is "
                                      "outside the scope of colorvideovdp_b200")
        # a user-defined photometry: run its forward(), then the colour matrix on the device
        I_lin = self.forward(I_src)
        if I_lin.shape[-4] != 3:
            return I_lin
        return self.linear_2_target_colorspace(I_lin, target_colorspace)

    def linear_2_target_colorspace(self, RGB_lin, target_colorspace):
        """pycvvdp/display_model.py:241-276: 3x3 colour matrix applied by the CUDA front end."""
        if target_colorspace not in _COLORSPACE_IDS:
            raise RuntimeError(f"Unknown colorspace '{target_colorspace}'")
        lin = vvdp_display_photo_eotf(1.0, contrast=1.0, EOTF="linear", exposure=1)
        lin.rgb2xyz_list = self.rgb2xyz_list
        # identity photometry: clip range wide open, no black level / reflection
        return lin._run_frontend(RGB_lin, _COLORSPACE_IDS[target_colorspace], passthrough=True)


class vvdp_display_photo_eotf(vvdp_display_photometry):
    """Display with an EOTF (sRGB, PQ, HLG, linear or a numeric gamma), pycvvdp/display_model.py:278-388."""

    def __init__(self, Y_peak, contrast=1000, source_colorspace="sRGB", EOTF=None, E_ambient=0, k_refl=0.005,
                 exposure=1, name=None, config_paths=[]):
        super().__init__(source_colorspace=source_colorspace, config_paths=config_paths)
        if EOTF is not None:
            self.EOTF = EOTF
        self.Y_peak = Y_peak
        self.contrast = contrast
        self.E_ambient = E_ambient
        self.k_refl = k_refl
        self.name = name
        self.exposure = exposure
        self._ctx = None

    def is_input_display_encoded(self):
        return self.EOTF != "linear"

    def __eq__(self, other):
        if not isinstance(other, self.__class__):
            return NotImplemented
        return (self.Y_peak == other.Y_peak and self.contrast == other.contrast and self.EOTF == other.EOTF
                and self.E_ambient == other.E_ambient and self.k_refl == other.k_refl
                and self.exposure == other.exposure)

    __hash__ = object.__hash__

    def get_peak_luminance(self):
        return self.Y_peak

    def get_black_level(self):
        Y_refl = self.E_ambient / math.pi * self.k_refl
        Y_black = self.Y_peak / self.contrast
        return Y_black, Y_refl

    def print(self):
        Y_black, Y_refl = self.get_black_level()
        logging.info("Photometric display model: {}".format(self.name))
        logging.info("  Peak luminance: {} cd/m^2".format(self.Y_peak))
        logging.info("  EOTF: {}".format(self.EOTF))
        logging.info("  Contrast - theoretical: {}:1".format(round(self.contrast)))
        logging.info("  Contrast - effective: {}:1".format(round(self.Y_peak / (Y_black + Y_refl))))
        logging.info("  Ambient light: {} lux".format(self.E_ambient))
        logging.info("  Display reflectivity: {}%".format(self.k_refl * 100))

    # ---- native description -------------------------------------------------------------------
    def native_display(self, ppd=1.0, passthrough=False) -> N.Display:
        """The cvvdp_b200_display struct of this display (EOTF id, photometry, RGB->XYZ)."""
        d = N.Display()
        if self.EOTF in _EOTF_IDS:
            d.eotf = _EOTF_IDS[self.EOTF]
            d.gamma = 1.0
            if self.EOTF == "HLG":  # display_model.py:351-355
                gamma = 1.2
                if self.Y_peak > 1000:
                    gamma = 1.2 + 0.42 * math.log10(self.Y_peak / 1000) - 0.07623 * math.log10(self.E_ambient / 5)
                d.gamma = gamma
        elif self.EOTF[0].isnumeric():
            d.eotf = N.EOTF_GAMMA
            d.gamma = float(self.EOTF)
        else:
            raise RuntimeError(f"Unknown EOTF '{self.EOTF}'")
        if passthrough:  # used by linear_2_target_colorspace: values go through unchanged
            d.eotf, d.Y_peak, d.contrast, d.E_ambient, d.k_refl, d.exposure = N.EOTF_NONE, 1.0, 1.0, 0.0, 0.0, 1.0
        else:
            d.Y_peak, d.contrast = float(self.Y_peak), float(self.contrast)
            d.E_ambient, d.k_refl, d.exposure = float(self.E_ambient), float(self.k_refl), float(self.exposure)
        rgb2xyz = getattr(self, "rgb2xyz_list", None)
        if rgb2xyz is None:  # 'luminance' colour space: no primaries, single-channel content only
            rgb2xyz = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
        for i in range(3):
            for j in range(3):
                d.rgb2xyz[i * 3 + j] = float(rgb2xyz[i][j])
        d.ppd = float(ppd)
        return d

    def _frontend_ctx(self, device):
        from . import cvvdp_metric as cm  # lazy: avoids an import cycle
        mock = cm._mock_library  # tests only (mock device)
        key = 0 if mock is not None else (device.index if device.index is not None else torch.cuda.current_device())
        if self._ctx is None or self._ctx[0] != (key, mock):
            params, lut = cm._default_native_inputs()
            self._ctx = ((key, mock), N.Context(params, lut, key, library=mock))
        return self._ctx[1]

    def _run_frontend(self, V, colorspace_id, passthrough=False):
        from . import cvvdp_metric as cm
        mock = cm._mock_library is not None
        if not mock and not torch.cuda.is_available():
            raise RuntimeError("colorvideovdp_b200 needs a CUDA device (no CPU fallback)")
        if V.dim() != 5:
            raise RuntimeError("expected a [B,C,1,H,W] frame")
        if not mock and not V.is_cuda:
            V = V.to(torch.device("cuda", torch.cuda.current_device()))  # host frames go to the caller's current device
        if V.dtype != torch.float32:
            V = V.to(torch.float32)
        B, Cc, F, H, W = V.shape
        ctx = self._frontend_ctx(V.device)
        ctx.set_display(self.native_display(passthrough=passthrough))
        clip = N.Clip()
        clip.data = V.data_ptr()
        for i, s in enumerate(V.stride()):
            clip.stride[i] = s
        clip.frame0, clip.n_frames = 0, F
        flags = torch.zeros(3, dtype=torch.int32, device=V.device)
        out = torch.empty((B, Cc, F, H, W), dtype=torch.float32, device=V.device)
        stream = None if mock else torch.cuda.current_stream(V.device).cuda_stream
        for f in range(F):
            dst = out[:, :, f]
            if not dst.is_contiguous():  # F > 1 only
                tmp = torch.empty((B, Cc, H, W), dtype=torch.float32, device=V.device)
                ctx.frontend(clip, B, Cc, H, W, N.DTYPE_F32, f, colorspace_id, tmp.data_ptr(), flags.data_ptr(), stream)
                out[:, :, f] = tmp
            else:
                ctx.frontend(clip, B, Cc, H, W, N.DTYPE_F32, f, colorspace_id, dst.data_ptr(), flags.data_ptr(), stream)
        if not passthrough and self.EOTF != "linear" and int(flags[0]) > 0:
            logging.warning("Pixel outside the valid range 0-1")  # display_model.py:335-337
        return out

    def forward(self, V):
        """Display-encoded values -> absolute linear light (cd/m^2), pycvvdp/display_model.py:333-365."""
        return self._run_frontend(V, N.CS_RGB_LINEAR)


class vvdp_display_geometry:
    """Effective resolution in pixels per degree (pycvvdp/display_model.py:431-626)."""

    def __init__(self, resolution, distance_m=None, distance_display_heights=None, fov_horizontal=None,
                 fov_vertical=None, fov_diagonal=None, diagonal_size_inches=None, ppd=None):
        self.resolution = resolution
        ar = resolution[0] / resolution[1]
        self.fixed_ppd = ppd
        if ppd is not None:
            return
        if diagonal_size_inches is not None:
            height_mm = math.sqrt((diagonal_size_inches * 25.4) ** 2 / (1 + ar ** 2))
            self.display_size_m = (ar * height_mm / 1000, height_mm / 1000)
        if distance_m is not None and distance_display_heights is not None:
            raise RuntimeError("You can pass only one of: 'distance_m', 'distance_display_heights'.")
        fovs = [fov_horizontal, fov_vertical, fov_diagonal]
        if distance_m is not None:
            self.distance_m = distance_m
        elif distance_display_heights is not None:
            if not hasattr(self, "display_size_m"):
                raise RuntimeError("You need to specify display diagonal size 'diagonal_size_inches' to specify "
                                   "viewing distance as 'distance_display_heights'")
            self.distance_m = distance_display_heights * self.display_size_m[1]
        elif any(f is not None for f in fovs):
            self.distance_m = 3  # default viewing distance for VR headsets
        else:
            raise RuntimeError("Viewing distance must be specified as 'distance_m' or 'distance_display_heights'.")
        if sum(f is not None for f in fovs) > 1:
            raise RuntimeError("You can pass only one of 'fov_horizontal', 'fov_vertical', 'fov_diagonal'. The other "
                               "dimensions are inferred from the resolution assuming that the pixels are square.")
        if fov_horizontal is not None:
            width_m = 2 * math.tan(math.radians(fov_horizontal / 2)) * self.distance_m
            self.display_size_m = (width_m, width_m / ar)
        elif fov_vertical is not None:
            height_m = 2 * math.tan(math.radians(fov_vertical / 2)) * self.distance_m
            self.display_size_m = (height_m * ar, height_m)
        elif fov_diagonal is not None:
            distance_px = math.hypot(resolution[0], resolution[1]) / (2.0 * math.tan(math.radians(fov_diagonal * 0.5)))
            height_deg = math.degrees(math.atan(resolution[1] / 2 / distance_px)) * 2
            height_m = 2 * math.tan(math.radians(height_deg / 2)) * self.distance_m
            self.display_size_m = (height_m * ar, height_m)
        self.display_size_deg = tuple(2 * math.degrees(math.atan(s / (2 * self.distance_m)))
                                      for s in self.display_size_m)

    def __eq__(self, other):
        if not isinstance(other, self.__class__):
            return NotImplemented
        return (self.resolution == other.resolution and self.distance_m == other.distance_m
                and self.display_size_m == other.display_size_m)

    __hash__ = object.__hash__

    def get_ppd(self, eccentricity=None):
        if self.fixed_ppd is not None:
            return self.fixed_ppd
        pix_deg = 2 * math.degrees(math.atan(0.5 * self.display_size_m[0] / self.resolution[0] / self.distance_m))
        base_ppd = 1 / pix_deg
        if eccentricity is None:
            return base_ppd
        delta = pix_deg / 2
        tan_delta = math.tan(math.radians(delta))
        tan_a = torch.tan(torch.deg2rad(eccentricity))
        return base_ppd * (torch.tan(torch.deg2rad(eccentricity + delta)) - tan_a) / tan_delta

    def print(self):
        logging.info("Geometric display model:")
        if self.fixed_ppd is not None:
            logging.info("  Fixed pixels-per-degree: {}".format(self.fixed_ppd))
        else:
            logging.info("  Resolution: {w} x {h} pixels".format(w=self.resolution[0], h=self.resolution[1]))
            logging.info("  Display size: {w:.1f} x {h:.1f} cm".format(w=self.display_size_m[0] * 100,
                                                                      h=self.display_size_m[1] * 100))
            logging.info("  Display size: {w:.2f} x {h:.2f} deg".format(w=self.display_size_deg[0],
                                                                       h=self.display_size_deg[1]))
            logging.info("  Viewing distance: {d:.3f} m".format(d=self.distance_m))
            logging.info("  Pixels-per-degree (center): {ppd:.2f}".format(ppd=self.get_ppd()))

    @classmethod
    def load(cls, display_name, config_paths=[]):
        models_file = utils.config_files.find("display_models.json", config_paths)
        models = utils.json2dict(models_file)
        if display_name not in models:
            logging.error(f"Display model: '{display_name}' not found in '{models_file}'")
            raise RuntimeError("Display model not found")
        model = models[display_name]
        assert "resolution" in model
        W, H = model["resolution"]
        if "pixels_per_degree" in model:
            return vvdp_display_geometry((W, H), ppd=model["pixels_per_degree"])
        inch = 0.0254
        if "viewing_distance_meters" in model:
            distance_m = model["viewing_distance_meters"]
        elif "viewing_distance_inches" in model:
            distance_m = model["viewing_distance_inches"] * inch
        else:
            distance_m = None
        if "diagonal_size_meters" in model:
            diag = model["diagonal_size_meters"] / inch
        else:
            diag = model.get("diagonal_size_inches")
        return vvdp_display_geometry((W, H), distance_m=distance_m, fov_diagonal=model.get("fov_diagonal"),
                                     diagonal_size_inches=diag)
