"""CPU oracle of the occupancy-grid update.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, smoke() and bench.py's
CPU legs; never by the product.

Restates, with explicit state instead of the reference's function attribute
(``update_occupancy_map.occupancy_probs``, process.py:122-125):

  * ``bresenham_line``               duc/ICP_LIDAR/process.py:86-112
  * ``update_occupancy_map``         duc/ICP_LIDAR/process.py:114-177
                                     (same body: duc/ICP_LIDAR/slam_offline.py:174-236)
  * ``filter_points_by_occupancy``   duc/ICP_LIDAR/process.py:203-226 (``filter_new_points_by_occupancy``)
                                     and :228-249 (``prune_global_map``: same predicate)

Parity status: PINNED.  ``tests/golden/make_golden_occupancy.py`` imports the unmodified
``process.py`` (with a stub ``open3d``) in the build container and records its outputs as
``tests/golden/reference_occupancy_golden.npz``; ``tests/test_occupancy_oracle.py`` checks this
restatement against them bit for bit.

Scalar arithmetic follows the installed NumPy 2 (NEP 50): ``np.float32 * 0.9`` is a float32
product with ``float32(0.9)``, ``occ >= 0.65`` compares with ``float32(0.65)``.  (Under
NumPy 1.x the same source promoted the scalar to float64 and rounded on the store; the
reference pins no NumPy version, so the installed one defines the behaviour, as for SciPy in
oracle/icp_oracle.py.)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

THRESHOLD_UP = 0.65          # process.py:158 (threshold_down, :157, is unused by live code)


def bresenham_line(x0, y0, x1, y1):
    """process.py:86-112, verbatim behaviour: cells from (x0, y0) to (x1, y1) inclusive."""
    cells = []
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    x, y = x0, y0
    sx = -1 if x0 > x1 else 1
    sy = -1 if y0 > y1 else 1
    if dx > dy:
        err = dx / 2.0
        while x != x1:
            cells.append((x, y))
            err -= dy
            if err < 0:
                y += sy
                err += dx
            x += sx
    else:
        err = dy / 2.0
        while y != y1:
            cells.append((x, y))
            err -= dx
            if err < 0:
                x += sx
                err += dy
            y += sy
    cells.append((x1, y1))
    return cells


def bresenham_cell(x0, y0, x1, y1, k):
    """Closed form of cell ``k`` of :func:`bresenham_line` (what the device kernel evaluates):
    with doubled integers the error term stays in [0, 2*major), so the minor axis has advanced
    ``floor((2*k*minor + major - 1) / (2*major))`` times after ``k`` steps."""
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx = -1 if x0 > x1 else 1
    sy = -1 if y0 > y1 else 1
    if dx > dy:
        return x0 + sx * k, y0 + sy * ((2 * k * dy + dx - 1) // (2 * dx))
    if dy == 0:
        return x0, y0
    return x0 + sx * ((2 * k * dx + dy - 1) // (2 * dy)), y0 + sy * k


def window_of(h, w, robot_x_px, robot_y_px, area):
    """The slice ``[y1:y2, x1:x2]`` of process.py:129-136 with Python's slice semantics
    (a negative stop counts from the end).  Returns (x1, y1, width, height)."""
    x1 = max(0, robot_x_px - area)
    y1 = max(0, robot_y_px - area)
    x2 = min(w, robot_x_px + area)
    y2 = min(h, robot_y_px + area)
    width = len(range(*slice(x1, x2).indices(w)))
    height = len(range(*slice(y1, y2).indices(h)))
    return x1, y1, width, height


def update_occupancy_map(occ, image, points_global, robot_pos, map_center_px, resolution,
                         p_occ_inc=0.2, p_free_dec=0.9, area=140):
    """process.py:114-177 on explicit state: ``occ`` (h, w) float32 probabilities and ``image``
    (h, w, 3) uint8, both updated in place."""
    if len(points_global) == 0:                                   # :116-117
        return
    h, w = image.shape[:2]
    robot_x_px = int(map_center_px[0] + robot_pos[0] / resolution)   # :128
    robot_y_px = int(map_center_px[1] - robot_pos[1] / resolution)   # :129
    x1 = max(0, robot_x_px - area)
    y1 = max(0, robot_y_px - area)
    x2 = min(w, robot_x_px + area)
    y2 = min(h, robot_y_px + area)
    image_new = image[y1:y2, x1:x2, :]
    occ_new = occ[y1:y2, x1:x2]
    height, width = image_new.shape[:2]
    rx, ry = robot_x_px - x1, robot_y_px - y1
    for pt in points_global:                                      # :146
        px = int(map_center_px[0] + pt[0] / resolution - x1)
        py = int(map_center_px[1] - pt[1] / resolution - y1)
        if not (0 <= px < width and 0 <= py < height):
            continue
        line = bresenham_line(rx, ry, px, py)
        last = len(line) - 1
        for i, (x, y) in enumerate(line):
            if not (0 <= x < width and 0 <= y < height):
                continue
            if i == last:
                occ_new[y, x] = min(1.0, occ_new[y, x] + p_occ_inc)    # :163
            else:
                if occ_new[y, x] >= THRESHOLD_UP:                      # :165-166
                    break
                occ_new[y, x] = max(0.0, occ_new[y, x] * p_free_dec)   # :167
    occ_uint8 = ((1 - occ_new) * 255).astype(np.uint8)            # :172
    image_new[:, :, 0] = occ_uint8
    image_new[:, :, 1] = occ_uint8
    image_new[:, :, 2] = occ_uint8


def filter_points_by_occupancy(points, occ, map_center_px, resolution, free_threshold=0.2):
    """process.py:203-226 / :228-249: drop the points that fall into a cell whose probability is
    below ``free_threshold``; points outside the grid are kept.  Returns the kept indices."""
    height, width = occ.shape
    keep = []
    for i, point in enumerate(points):
        px = int(map_center_px[0] + point[0] / resolution)
        py = int(map_center_px[1] - point[1] / resolution)
        if not (0 <= px < width and 0 <= py < height):
            keep.append(i)
            continue
        if occ[py, px] < free_threshold:
            continue
        keep.append(i)
    return np.asarray(keep, dtype=np.int64)


# ---- plain-C restatement (oracle/occupancy_oracle.c) for long replays ------------------------
_SO = os.path.join(_HERE, "_build", "libocc_oracle.so")
_lib = None


def _clib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "occupancy_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
    return _lib


def update_occupancy_map_c(occ, image, points_global, robot_pos, map_center_px, resolution,
                           p_occ_inc=0.2, p_free_dec=0.9, area=140):
    """Same contract as :func:`update_occupancy_map`, in C (float32 arithmetic spelled out)."""
    assert occ.dtype == np.float32 and occ.flags.c_contiguous
    assert image.dtype == np.uint8 and image.flags.c_contiguous and image.shape[2] == 3
    if len(points_global) == 0:
        return
    pts = np.ascontiguousarray(np.asarray(points_global, dtype=np.float64)[:, :2])
    h, w = occ.shape
    _clib().occ_update(occ.ctypes.data_as(C.c_void_p), image.ctypes.data_as(C.c_void_p),
                       C.c_int(h), C.c_int(w), pts.ctypes.data_as(C.c_void_p), C.c_int(len(pts)),
                       C.c_double(robot_pos[0]), C.c_double(robot_pos[1]),
                       C.c_double(map_center_px[0]), C.c_double(map_center_px[1]),
                       C.c_double(resolution), C.c_float(np.float32(p_occ_inc)),
                       C.c_float(np.float32(p_free_dec)), C.c_int(area))


def synth_replay(seed, frames, beams=180, room=3500.0, step=60.0):
    """Seeded replay for tests/bench: a robot random-walking inside a star-convex room (the room
    family of icp_oracle.synth_room_batch), ``beams`` returns per frame in the MAP frame.
    Returns (points [frames, beams, 2] float64, robot_xy [frames, 2] float64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.uniform(0.0, 0.15 * room, size=4)
    psi = rng.uniform(0.0, 2.0 * np.pi, size=4)
    pos = np.zeros(2)
    pts = np.empty((frames, beams, 2))
    rob = np.empty((frames, 2))
    heading = 0.0
    for f in range(frames):
        heading += rng.normal(0.0, 0.3)
        nxt = pos + step * np.array([np.cos(heading), np.sin(heading)])
        if np.hypot(*nxt) < 0.45 * room:
            pos = nxt
        else:
            heading += np.pi
        phi = np.sort(rng.uniform(0.0, 2.0 * np.pi, size=beams))
        wall = room + (a[None, :] * np.sin(np.arange(1, 5)[None, :] * phi[:, None] + psi[None, :])).sum(1)
        wall = wall + rng.normal(0.0, 5.0, size=beams)
        # wall point in the room frame, seen from pos: the beam ends on the wall along phi from the origin
        pts[f, :, 0] = wall * np.cos(phi)
        pts[f, :, 1] = wall * np.sin(phi)
        rob[f] = pos
    return pts, rob


def unpack_scan(packed, f):
    """Raw polar rows of scan ``f`` (0-based) of tests/golden/scan_data_1_packed.npz."""
    off = packed["offsets"]
    a, b = int(off[f]), int(off[f + 1])
    return np.stack([packed["quality"][a:b].astype(np.float64),
                     packed["angle64"][a:b].astype(np.float64) / 64.0,
                     packed["dist4"][a:b].astype(np.float64) / 4.0], axis=1)


def replay_frame(packed, poses, f):
    """Inputs of frame ``f`` of the recording replay: the scan's Cartesian points
    (process.py:38-52 restatement) moved into the map frame by ``poses[f] = (cos, sin, tx, ty)``
    with element-wise float64 operations (no BLAS, so every host derives the same bits), as
    (points [n, 3], robot_pos [3])."""
    from . import icp_oracle as orc
    xy = orc.polar_to_cartesian(unpack_scan(packed, f))
    c, s, tx, ty = (float(v) for v in poses[f])
    pts = np.zeros((len(xy), 3))
    if len(xy):
        pts[:, 0] = c * xy[:, 0] - s * xy[:, 1] + tx
        pts[:, 1] = s * xy[:, 0] + c * xy[:, 1] + ty
    return pts, np.array([tx, ty, 0.0])
