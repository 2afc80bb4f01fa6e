"""colorvideovdp_b200 -- B200-native engine for the ColorVideoVDP hot path.

Drop-in for the `pycvvdp` names on that path: ``cvvdp`` (metric), ``vq_metric`` / ``register_metric``
(plugin API), ``video_source*`` and ``vvdp_display_*`` (plugin surfaces).  The arithmetic lives in the
CUDA library built from ``csrc/`` (see include/cvvdp_b200.h); importing the package does not need a
GPU, constructing a metric does.
"""
from .vq_metric import vq_metric, vq_exception, vq_metric_dict, register_metric
from .video_source import video_source, video_source_dm, video_source_array, reshuffle_dims
from .display_model import vvdp_display_photometry, vvdp_display_photo_eotf, vvdp_display_geometry
from .cvvdp_metric import cvvdp
from .video_source_yuv import video_source_yuv_file, YUVReader, video_reader_yuv, decode_video_props, create_yuv_fname
from .video_source_file import (video_source_video_file, video_source_temp_resample_file, video_reader,
                                video_reader_yuv_pytorch)
from .utils import config_files

__version__ = "0.1.0"
