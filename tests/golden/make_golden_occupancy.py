#!/usr/bin/env python
"""Generate tests/golden/reference_occupancy_golden.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_occupancy.py

The reference's ``duc/ICP_LIDAR/process.py`` is imported as it is (oracle/ref_loader.py puts an
empty ``open3d`` stub in place for the import) and these functions are called:

  bresenham_line                    process.py:86-112    every end point of a 19 x 19 block from two
                                                          start cells + 200 seeded long lines
  update_occupancy_map              process.py:114-177   (a) replay of the first 200 scans of the
                                                          bundled Scan_data_1 recording on the
                                                          reference's own map geometry (Config.py:
                                                          833 x 1000 cells, 30 mm), poses = chain of
                                                          the reference icp() results of
                                                          reference_icp_golden.npz; (b) 48 seeded
                                                          small cases with preset probabilities
                                                          (exact float32(0.65) cells, saturated and
                                                          denormal cells), robots on the border and
                                                          outside the map, empty inputs, non-default
                                                          parameters
  filter_new_points_by_occupancy    process.py:203-226   seeded points over preset grids
  polar_to_cartesian_3d             process.py:38-52     all 1,831 scans of Scan_data_1: point count and
                                                          CRC32 of the float64 output

The inputs of (b) and of the filter cases are stored with the outputs; the inputs of (a) are
re-derived by the tests from scan_data_1_packed.npz and the stored per-frame poses with
element-wise float64 operations only (no BLAS), so the fixture stays small.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import icp_oracle as orc          # noqa: E402
from oracle import occupancy_oracle as occ_orc  # noqa: E402
from oracle import ref_loader                  # noqa: E402

REPLAY_FRAMES = 200
CRC_FRAMES = [0, 1, 2, 10, 50, 100, 150, 199]
MAP_H, MAP_W, RES = 833, 1000, 30              # Config.py:7-9,25-26
CENTER = (MAP_W // 2, MAP_H // 2)              # slam_offline.py:320


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def ref_update(ref, occ, image, pts, robot, center, res, **kw):
    """Call the reference with explicit state: its probabilities live in a function attribute."""
    ref.update_occupancy_map.occupancy_probs = occ
    ref.update_occupancy_map(image, pts, robot, center, res, **kw)
    return ref.update_occupancy_map.occupancy_probs


def small_cases(rng):
    """Seeded inputs of the small cases (shared with the tests through the fixture)."""
    special = np.array([0.65, 0.64999, 0.7, 1.0, 0.0, 0.5, 0.2, 0.45, 1e-44, 0.95], dtype=np.float32)
    cases = []
    for c in range(48):
        h, w = int(rng.integers(30, 70)), int(rng.integers(30, 90))
        area = int(rng.integers(5, 40))
        res = [30, 30, 25.0, 50][c % 4]
        center = (w // 2, h // 2)
        occ = np.full((h, w), 0.5, dtype=np.float32)
        k = int(rng.integers(0, h * w // 3))
        occ.reshape(-1)[rng.integers(0, h * w, size=k)] = special[rng.integers(0, len(special), size=k)]
        image = np.full((h, w, 3), 128, dtype=np.uint8)
        mode = c % 8
        robot = rng.uniform(-0.4, 0.4, size=3) * np.array([w * res, h * res, 0.0])
        if mode == 1:                 # robot on the border
            robot[0] = -center[0] * res + rng.uniform(0, 2 * res)
        elif mode == 2:               # robot outside the map
            robot[:2] = np.array([w, h]) * res * rng.uniform(0.55, 0.9) * rng.choice([-1, 1], size=2)
        elif mode == 3:               # robot in the top-left corner cell region
            robot[:2] = np.array([-center[0] * res + 1.0, center[1] * res - 1.0])
        n = 0 if c == 5 else int(rng.integers(1, 60))
        cols = 2 if c % 3 == 0 else 3
        pts = np.zeros((n, cols))
        pts[:, :2] = robot[:2] + rng.uniform(-1.3, 1.3, size=(n, 2)) * area * res
        if n > 3:                     # the ray that ends in the robot cell and exact cell edges
            pts[0, :2] = robot[:2]
            pts[1, :2] = np.round(pts[1, :2] / res) * res
        kw = {}
        if c % 6 == 4:
            kw = dict(p_occ_inc=0.35, p_free_dec=0.8)
        cases.append(dict(occ=occ, image=image, pts=pts, robot=robot, center=center, res=res, area=area, kw=kw))
    return cases


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    ref = ref_loader.load_reference_process()
    out = {}

    # ---- bresenham_line -----------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(86))
    ends = [(x0, y0, x0 + dx, y0 + dy) for (x0, y0) in ((0, 0), (3, -2))
            for dx in range(-9, 10) for dy in range(-9, 10)]
    ends += [tuple(int(v) for v in rng.integers(-300, 300, size=4)) for _ in range(200)]
    cells, offs = [], [0]
    for e in ends:
        line = ref.bresenham_line(*e)
        cells.extend(line)
        offs.append(len(cells))
    out["bres_ends"] = np.asarray(ends, dtype=np.int32)
    out["bres_offsets"] = np.asarray(offs, dtype=np.int32)
    out["bres_cells"] = np.asarray(cells, dtype=np.int16)

    # ---- replay of the recording ----------------------------------------------------------
    gold = np.load(os.path.join(HERE, "reference_icp_golden.npz"))
    packed = dict(np.load(os.path.join(HERE, "scan_data_1_packed.npz")))
    th, tt = gold["pair_theta_tot"], gold["pair_t_tot"]
    poses = np.zeros((REPLAY_FRAMES, 4))          # cos, sin, tx, ty of scan f+1 in the map frame
    c, s, tx, ty = 1.0, 0.0, 0.0, 0.0
    for f in range(REPLAY_FRAMES):
        poses[f] = (c, s, tx, ty)
        pc, ps, px, py = np.cos(th[f]), np.sin(th[f]), tt[f, 0], tt[f, 1]   # scan f+2 -> scan f+1
        c, s, tx, ty = c * pc - s * ps, s * pc + c * ps, c * px - s * py + tx, s * px + c * py + ty
    out["replay_poses"] = poses
    occ = None
    image = np.full((MAP_H, MAP_W, 3), 128, dtype=np.uint8)
    if hasattr(ref.update_occupancy_map, "occupancy_probs"):
        del ref.update_occupancy_map.occupancy_probs          # first call creates it (process.py:122-123)
    crcs = []
    for f in range(REPLAY_FRAMES):
        pts, robot = orc_replay_frame(packed, poses, f)
        if len(pts) == 0:
            continue
        ref.update_occupancy_map(image, pts, robot, CENTER, RES)
        occ = ref.update_occupancy_map.occupancy_probs
        if f in CRC_FRAMES:
            crcs.append((f, crc(occ), crc(image)))
    out["replay_crc"] = np.asarray(crcs, dtype=np.int64)
    out["replay_occ_final"] = occ
    out["replay_image_final"] = image[:, :, 0].copy()
    assert np.array_equal(image[:, :, 0], image[:, :, 1]) and np.array_equal(image[:, :, 0], image[:, :, 2])

    # ---- small seeded cases ---------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(114))
    cases = small_cases(rng)
    out["small_count"] = np.int32(len(cases))
    for i, cs in enumerate(cases):
        occ_in, img_in = cs["occ"].copy(), cs["image"].copy()
        occ_out = ref_update(ref, cs["occ"], cs["image"], cs["pts"], cs["robot"], cs["center"], cs["res"],
                             area=cs["area"], **cs["kw"])
        out[f"small_{i}_occ_in"] = occ_in
        out[f"small_{i}_pts"] = cs["pts"]
        out[f"small_{i}_robot"] = cs["robot"]
        out[f"small_{i}_par"] = np.asarray([cs["center"][0], cs["center"][1], cs["res"], cs["area"],
                                            cs["kw"].get("p_occ_inc", 0.2), cs["kw"].get("p_free_dec", 0.9)])
        out[f"small_{i}_occ_out"] = occ_out
        out[f"small_{i}_img_out"] = cs["image"][:, :, 0].copy()
        assert img_in.shape == cs["image"].shape

    # ---- filter_new_points_by_occupancy ---------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(203))
    for i in range(6):
        h, w = int(rng.integers(40, 90)), int(rng.integers(40, 90))
        res = [30, 25.0, 40][i % 3]
        center = (w // 2, h // 2)
        grid = rng.choice(np.array([0.2, 0.19999999, 0.0, 0.5, 0.7, 0.1, 0.20000002], dtype=np.float32), size=(h, w))
        pts = rng.uniform(-0.7, 0.7, size=(400, 3)) * np.array([w * res, h * res, 0.0])
        kept = ref.filter_new_points_by_occupancy(pts, grid, center, res)
        thr = {}
        if i == 4:
            thr = dict(free_threshold=0.5)
            kept = ref.filter_new_points_by_occupancy(pts, grid, center, res, **thr)
        out[f"filter_{i}_grid"] = grid
        out[f"filter_{i}_pts"] = pts
        out[f"filter_{i}_par"] = np.asarray([center[0], center[1], res, thr.get("free_threshold", 0.2)])
        out[f"filter_{i}_kept"] = kept
    out["filter_count"] = np.int32(6)

    # ---- polar_to_cartesian_3d (process.py:38-52) on every scan of the recording --------------
    # (the registration goldens were produced through the oracle's restatement of this function;
    # here the unmodified function itself is recorded: point count and CRC32 of its output bytes)
    n_scans = len(packed["offsets"]) - 1
    p2c_n = np.zeros(n_scans, dtype=np.int32)
    p2c_crc = np.zeros(n_scans, dtype=np.uint32)
    for f in range(n_scans):
        pts = ref.polar_to_cartesian_3d(occ_orc.unpack_scan(packed, f))
        p2c_n[f] = len(pts)
        p2c_crc[f] = crc(np.asarray(pts, dtype=np.float64)) if len(pts) else 0
    out["p2c_count"] = p2c_n
    out["p2c_crc32"] = p2c_crc

    path = os.path.join(HERE, "reference_occupancy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def orc_replay_frame(packed, poses, f):
    return occ_orc.replay_frame(packed, poses, f)


if __name__ == "__main__":
    main()
