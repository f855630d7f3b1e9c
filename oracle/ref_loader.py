"""Import the UNMODIFIED reference ``labels_segmentation/icp.py``.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container, never on the GPU
box.  ``make -C oracle ref`` (run by ``__graft_entry__.build()`` when the reference tree is
present) places a byte-identical copy of that one file at ``oracle/_ref/icp.py``; the directory is
git-ignored (no reference source enters the history) but travels to the GPU box with the other
built artefacts, so that ``bench.py``'s CPU arm can time the reference itself there.  Callers must
check :func:`reference_available` first.
The reference module imports ``matplotlib.pyplot`` (icp.py:2) and runs a demo
with ``plt.show()`` at import time (icp.py:55-78); matplotlib is not installed
here, so a no-op stub is registered for the duration of the import.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ICP_REFERENCE_ROOT", "/root/reference")
_ICP_TREE = os.path.join(REFERENCE_ROOT, "labels_segmentation", "icp.py")
_ICP_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "icp.py")
_ICP_PATH = _ICP_TREE if os.path.isfile(_ICP_TREE) else _ICP_COPY
_cached = None


def reference_available() -> bool:
    return os.path.isfile(_ICP_PATH)


def reference_icp_path() -> str:
    """Where the unmodified icp.py is loaded from (the tree, else the build-time copy)."""
    return _ICP_PATH


def scan_dir(name="Scan_data_1") -> str:
    return os.path.join(REFERENCE_ROOT, name)


def load_reference_icp():
    """Returns the reference module: ``.icp``, ``.best_fit_transform`` and the
    demo globals ``A, B, A_aligned, R_est, t_est`` (icp.py:57-67)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(_ICP_PATH)

    def _noop(*a, **k):
        return None

    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: _noop        # figure/scatter/legend/... -> no-op
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    try:
        spec = importlib.util.spec_from_file_location("_reference_icp", _ICP_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod


_PROCESS_PATH = os.path.join(REFERENCE_ROOT, "duc", "ICP_LIDAR", "process.py")
_cached_process = None


def load_reference_process():
    """The UNMODIFIED ``duc/ICP_LIDAR/process.py`` (build container only): ``bresenham_line``,
    ``update_occupancy_map``, ``filter_new_points_by_occupancy``, ``polar_to_cartesian_3d``.
    It imports ``open3d`` (process.py:1, and through gicp_lidar.py), which is not installed; an
    empty stub module stands in for the duration of the import (none of the functions named
    above touches it).  ``cv2``, ``Config`` and ``gicp_lidar`` are the real ones."""
    global _cached_process
    if _cached_process is not None:
        return _cached_process
    if not os.path.isfile(_PROCESS_PATH):
        raise FileNotFoundError(_PROCESS_PATH)
    ref_dir = os.path.dirname(_PROCESS_PATH)
    saved = {k: sys.modules.get(k) for k in ("open3d", "gicp_lidar", "Config")}
    sys.modules["open3d"] = types.ModuleType("open3d")
    sys.path.insert(0, ref_dir)
    try:
        spec = importlib.util.spec_from_file_location("_reference_process", _PROCESS_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(ref_dir)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_process = mod
    return mod
