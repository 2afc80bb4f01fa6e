"""Metric plugin API of the reference (pycvvdp/vq_metric.py): base class, exception, registry."""
import abc

from .video_source import video_source_array
from .display_model import vvdp_display_photometry


class vq_exception(Exception):
    def __init__(self, message):
        super().__init__(message)


class vq_metric:
    """Base class of video-quality metrics (pycvvdp/vq_metric.py:11-80)."""

    def predict(self, test_cont, reference_cont, dim_order="BCFHW", frames_per_second=0):
        test_vs = video_source_array(test_cont, reference_cont, frames_per_second, dim_order=dim_order,
                                     display_photometry=self.display_photometry)
        return self.predict_video_source(test_vs)

    @abc.abstractmethod
    def predict_video_source(self, vid_source):
        pass

    @abc.abstractmethod
    def quality_unit(self):
        pass

    def get_info_string(self):
        return None

    def set_display_model(self, display_name="standard_4k", display_photometry=None, display_geometry=None,
                          config_paths=[]):
        if display_photometry is None:
            self.display_photometry = vvdp_display_photometry.load(display_name, config_paths)
            self.display_name = display_name
        else:
            self.display_photometry = display_photometry
            self.display_name = "unspecified"

    def set_base_fname(self, base_fname):
        self.base_fname = base_fname

    def train(self, do_training=True):
        pass

    def short_name(self):
        return self.__class__.__name__.replace("_", "-")

    def export_distogram(self, stats, fname, jod_max=None, base_size=6):
        raise vq_exception(f"Metric {self.short_name()} cannot generate distograms")


vq_metric_dict = dict()


def register_metric(metric_class):
    vq_metric_dict[metric_class.__name__] = metric_class
