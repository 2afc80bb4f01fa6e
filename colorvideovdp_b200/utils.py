"""Configuration-file lookup shared by the display model, the CSF and the metric.

Mirrors the search order of the reference (pycvvdp/utils.py:133-174): explicit files in
``config_paths`` whose base name starts with the requested name, then directories in ``config_paths``,
then ``$CVVDP_PATH``, then the ``vvdp_data`` directory shipped with this package.
"""
import json
import os

_PKG_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vvdp_data")


def json2dict(file):
    if not os.path.isfile(file):
        raise RuntimeError(f"Error: Cannot find file {file}")
    with open(file, "r") as fh:
        return json.load(fh)


class config_files:
    @classmethod
    def find(cls, fname, config_paths):
        if not isinstance(config_paths, list):
            raise RuntimeError("config_paths must be a list")
        stem, ext = os.path.splitext(fname)
        for cp in config_paths:
            if not (os.path.isfile(cp) or os.path.isdir(cp)):
                raise RuntimeError(f"config_path '{cp}' does not exist")
        for cp in config_paths:
            base = os.path.basename(cp)
            if os.path.isfile(cp) and base.startswith(stem) and base.endswith(ext):
                return cp
        candidates = [os.path.join(cp, fname) for cp in config_paths if os.path.isdir(cp)]
        env_dir = os.getenv("CVVDP_PATH")
        if env_dir is not None:
            candidates.append(os.path.join(env_dir, fname))
        candidates.append(os.path.join(_PKG_DATA, fname))
        for path in candidates:
            if os.path.isfile(path):
                return path
        raise RuntimeError(f"The configuration file {fname} not found")
