"""Scan files -> device tables.

The on-disk format is what the recorder writes (duc/code python/read_lidar.py:72,137-138)
and ``load_and_prepare_scan`` validates (duc/ICP_LIDAR/process.py:16-22): one ``.npy`` per
scan, ``(N,3) float64`` rows ``[quality, angle_deg, distance_mm]`` (or ``(N,2)`` Cartesian).
File I/O and padding are host work; the numeric scan preparation (filter + polar ->
Cartesian, process.py:38-52) runs on the device via ``registration.polar_to_cartesian``.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .registration import ScanTable, polar_to_cartesian

_STEMS = ("Scan_data_", "scan_data_", "scan_")


def scan_file(directory: str, k: int) -> Optional[str]:
    """Resolve scan number k: the bundled recording mixes ``Scan_data_{k}.npy`` and
    ``scan_data_{k}.npy`` and the reference's config uses Windows separators
    (duc/ICP_LIDAR/Config.py:1)."""
    directory = directory.replace("\\", os.sep)
    for stem in _STEMS:
        p = os.path.join(directory, f"{stem}{k}.npy")
        if os.path.exists(p):
            return p
    return None


def load_raw_scan(path: Optional[str]) -> Optional[np.ndarray]:
    """np.load + the shape check of process.py:16-22; None on any failure
    (process.py:10-13,33-36 return None instead of raising)."""
    if path is None or not os.path.exists(path):
        return None
    try:
        a = np.load(path)
    except Exception:
        return None
    if a.ndim != 2 or a.shape[1] not in (2, 3):
        return None
    return np.asarray(a, dtype=np.float64)


def load_raw_sequence(directory: str, first: int, last: int) -> Tuple[List[np.ndarray], List[int]]:
    """Raw scans first..last (inclusive); unreadable files are skipped like the SLAM loop
    does (duc/ICP_LIDAR/slam_offline.py:346-349)."""
    scans, numbers = [], []
    for k in range(first, last + 1):
        a = load_raw_scan(scan_file(directory, k))
        if a is not None:
            scans.append(a)
            numbers.append(k)
    return scans, numbers


def unpack_fixture(npz_path: str) -> List[np.ndarray]:
    """Inverse of tests/golden/make_golden.py:pack_scans (bit-exact float64 rows)."""
    z = np.load(npz_path)
    off = z["offsets"]
    q = z["quality"].astype(np.float64)
    a = z["angle64"].astype(np.float64) / 64.0
    d = z["dist4"].astype(np.float64) / 4.0
    rows = np.stack([q, a, d], axis=1)
    return [rows[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def raw_table(scans: Sequence[np.ndarray], pitch: Optional[int] = None, pin: bool = False):
    """Pad polar scans to host tensors ([S,pitch,3] float64, [S] int32)."""
    longest = max((len(s) for s in scans), default=1)
    pitch = max(1, longest) if pitch is None else int(pitch)
    raw = np.zeros((len(scans), pitch, 3), dtype=np.float64)
    lens = np.zeros(len(scans), dtype=np.int32)
    for i, s in enumerate(scans):
        if s.shape[1] != 3:
            raise ValueError("raw_table takes polar (N,3) scans")
        raw[i, : len(s)] = s
        lens[i] = len(s)
    tr, tl = torch.from_numpy(raw), torch.from_numpy(lens)
    if pin:
        tr, tl = tr.pin_memory(), tl.pin_memory()
    return tr, tl


def prepare_scans(scans: Sequence[np.ndarray], device="cuda", out_pitch: Optional[int] = None,
                  filter="process") -> ScanTable:
    """Polar scans -> filtered Cartesian ScanTable, computed on the device.  ``filter``: which of
    the reference's copies of polar_to_cartesian_3d (registration.POLAR_FILTERS)."""
    raw, lens = raw_table(scans)
    return polar_to_cartesian(raw.to(device), lens.to(device), out_pitch=out_pitch, filter=filter)
