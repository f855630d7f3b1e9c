"""GPU parity of the occupancy-grid kernels (csrc/occupancy.cu, through the C ABI) against the
pinned oracle and against the reference's recorded outputs.  Bit-exact: float32 probabilities as
uint32 patterns, uint8 grey levels, kept rows."""
import os
import zlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import occupancy_oracle as occ     # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAP_H, MAP_W, RES = 833, 1000, 30              # Config.py:7-9,25-26
CENTER = (MAP_W // 2, MAP_H // 2)              # slam_offline.py:320


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import icp_slam_yolo_b200 as m
    m.lib()
    return m


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "reference_occupancy_golden.npz"))


@pytest.fixture(scope="module")
def packed():
    return dict(np.load(os.path.join(GOLDEN, "scan_data_1_packed.npz")))


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_small_reference_cases_bit_exact(pkg, gold):
    """The 48 seeded cases recorded from the unmodified reference: preset probabilities (exact
    float32(0.65), saturated, denormal cells), robots on the border / outside the map (Python's
    negative slice stop), empty input, non-default parameters, (N, 2) and (N, 3) points."""
    for i in range(int(gold["small_count"])):
        par = gold[f"small_{i}_par"]
        occ_in = gold[f"small_{i}_occ_in"]
        h, w = occ_in.shape
        grid = pkg.OccupancyGrid(h, w, (int(par[0]), int(par[1])), float(par[2]))
        grid.probs[0].copy_(torch.from_numpy(occ_in))
        grid.update(gold[f"small_{i}_pts"], gold[f"small_{i}_robot"], p_occ_inc=float(par[4]),
                    p_free_dec=float(par[5]), area=int(par[3]))
        assert np.array_equal(_bits(grid.probs_numpy()), _bits(gold[f"small_{i}_occ_out"])), f"case {i}"
        img = grid.image_numpy()
        for ch in range(3):
            assert np.array_equal(img[:, :, ch], gold[f"small_{i}_img_out"]), f"case {i} channel {ch}"


def test_recording_replay_frame_by_frame_and_single_launch(pkg, gold, packed):
    """200 frames of Scan_data_1 on the reference's map geometry (833 x 1000 cells, 30 mm):
    (a) one update() per frame, CRC32 of probabilities and picture at the recorded frames;
    (b) all frames in ONE launch (update_frames) -- same final arrays."""
    poses = gold["replay_poses"]
    want = {int(f): (int(a), int(b)) for f, a, b in gold["replay_crc"]}
    frames = [occ.replay_frame(packed, poses, f) for f in range(len(poses))]
    g1 = pkg.OccupancyGrid(MAP_H, MAP_W, CENTER, RES)
    for f, (pts, robot) in enumerate(frames):
        g1.update(pts, robot)
        if f in want:
            assert (_crc(g1.probs_numpy()), _crc(g1.image_numpy())) == want[f], f"frame {f}"
    assert np.array_equal(_bits(g1.probs_numpy()), _bits(gold["replay_occ_final"]))
    assert np.array_equal(g1.image_numpy()[:, :, 0], gold["replay_image_final"])

    pitch = max(len(p) for p, _ in frames)
    tab = np.zeros((1, len(frames), pitch, 2))
    lens = np.zeros((1, len(frames)), dtype=np.int32)
    rob = np.zeros((1, len(frames), 2))
    for f, (pts, robot) in enumerate(frames):
        tab[0, f, :len(pts)] = pts[:, :2]
        lens[0, f] = len(pts)
        rob[0, f] = robot[:2]
    g2 = pkg.OccupancyGrid(MAP_H, MAP_W, CENTER, RES)
    g2.update_frames(torch.from_numpy(tab).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(rob).cuda())
    assert np.array_equal(_bits(g2.probs_numpy()), _bits(gold["replay_occ_final"]))
    assert np.array_equal(g2.image_numpy()[:, :, 2], gold["replay_image_final"])


def test_batched_maps_against_c_oracle(pkg):
    """24 independent maps x 40 frames x 180 beams in one launch (one CTA per map), float32 and
    float64 point tables, against the C oracle (itself pinned to the reference's outputs)."""
    n_maps, frames, beams, h, w = 24, 40, 180, 360, 400
    center, res = (w // 2, h // 2), 30
    tab = np.zeros((n_maps, frames, beams, 2))
    rob = np.zeros((n_maps, frames, 2))
    for m in range(n_maps):
        tab[m], rob[m] = occ.synth_replay(100 + m, frames, beams=beams)
    lens = np.full((n_maps, frames), beams, dtype=np.int32)
    lens[3, 5] = 0
    lens[7, :] = 17
    for dtype in (np.float64, np.float32):
        t = tab.astype(dtype)
        grid = pkg.OccupancyGrid(h, w, center, res, n_maps=n_maps)
        grid.update_frames(torch.from_numpy(t).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(rob).cuda())
        probs = grid.probs.cpu().numpy()
        image = grid.image.cpu().numpy()
        for m in range(n_maps):
            o = np.full((h, w), 0.5, dtype=np.float32)
            im = np.full((h, w, 3), 128, dtype=np.uint8)
            for f in range(frames):
                occ.update_occupancy_map_c(o, im, t[m, f, :lens[m, f]].astype(np.float64), rob[m, f], center, res)
            assert np.array_equal(_bits(probs[m]), _bits(o)), f"map {m} {dtype.__name__}"
            assert np.array_equal(image[m], im), f"map {m} {dtype.__name__}"
        assert probs.max() >= 0.65 and probs.min() < 0.1


def test_large_window_loops_and_tiny_window(pkg):
    """area = 400 (rays longer than one 160-cell trip; cell lists leave room for few beams per
    chunk) and area = 0 / 1, against the C oracle."""
    h, w, res = 900, 900, 10
    pts, rob = occ.synth_replay(5, 12, beams=300)
    for area in (400, 1, 0):
        grid = pkg.OccupancyGrid(h, w, (450, 450), res)
        o = np.full((h, w), 0.5, dtype=np.float32)
        im = np.full((h, w, 3), 128, dtype=np.uint8)
        for f in range(len(pts)):
            grid.update(pts[f], rob[f], area=area)
            occ.update_occupancy_map_c(o, im, pts[f], rob[f], (450, 450), res, area=area)
        assert np.array_equal(_bits(grid.probs_numpy()), _bits(o)), f"area {area}"
        assert np.array_equal(grid.image_numpy(), im), f"area {area}"


def test_filter_points_reference_cases_and_large(pkg, gold):
    for i in range(int(gold["filter_count"])):
        par = gold[f"filter_{i}_par"]
        g = gold[f"filter_{i}_grid"]
        grid = pkg.OccupancyGrid(g.shape[0], g.shape[1], (int(par[0]), int(par[1])), float(par[2]), with_image=False)
        grid.probs[0].copy_(torch.from_numpy(g))
        kept = grid.filter_points(gold[f"filter_{i}_pts"], free_threshold=float(par[3]))
        assert np.array_equal(kept, gold[f"filter_{i}_kept"]), f"case {i}"
    # 300,000 map points (prune_global_map at scale), CUDA tensor in / out, (N, 2) float32 rows
    rng = np.random.Generator(np.random.PCG64(9))
    g = rng.choice(np.array([0.2, 0.19999999, 0.0, 0.5, 0.7], dtype=np.float32), size=(833, 1000))
    pts = (rng.uniform(-0.6, 0.6, size=(300_000, 2)) * np.array([1000 * 30, 833 * 30])).astype(np.float32)
    grid = pkg.OccupancyGrid(833, 1000, CENTER, RES, with_image=False)
    grid.probs[0].copy_(torch.from_numpy(g))
    kept = grid.filter_points(torch.from_numpy(pts).cuda()).cpu().numpy()
    want = occ.filter_points_by_occupancy(pts.astype(np.float64), g, CENTER, RES)
    assert np.array_equal(kept, pts[want])
    assert grid.filter_points(pts[:0]).shape == (0, 2)


def test_bad_arguments_raise(pkg):
    grid = pkg.OccupancyGrid(50, 60, (30, 25), 30)
    with pytest.raises(pkg.B200IcpError):
        grid.update_frames(torch.zeros(1, 1, 4, 2, dtype=torch.float64), None,
                           torch.zeros(1, 1, 2, dtype=torch.float64).cuda())          # CPU tensor
    with pytest.raises(pkg.B200IcpError):
        grid.update(np.zeros((3, 2)), (0.0, 0.0), area=20000)                           # window too large
    with pytest.raises(pkg.B200IcpError):
        pkg.OccupancyGrid(50, 60, (30, 25), 30, device="cpu")


def test_random_cases_arbitrary_probabilities(pkg):
    """300 seeded frames on small grids with arbitrary float32 probabilities (not only the values the
    update itself produces, a fifth of the cells exactly float32(0.65)), robots inside / on the
    border / outside the map, windows from 0 to 29 cells, random parameters: kernel == C oracle."""
    rng = np.random.Generator(np.random.PCG64(172))
    for q in range(300):
        h, w = int(rng.integers(8, 48)), int(rng.integers(8, 48))
        area = int(rng.integers(0, 30))
        res = float(rng.choice([10.0, 30.0, 37.5]))
        center = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        o = rng.random((h, w), dtype=np.float32)
        o[rng.random((h, w)) < 0.2] = np.float32(0.65)
        im = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        robot = np.array([rng.uniform(-1.5, 1.5) * w * res, rng.uniform(-1.5, 1.5) * h * res, 0.0])
        pts = np.zeros((int(rng.integers(0, 25)), 3))
        pts[:, :2] = robot[:2] + rng.normal(0, (area + 2) * res, size=(len(pts), 2))
        kw = dict(p_occ_inc=float(rng.choice([0.2, 0.35, 1.5])), p_free_dec=float(rng.choice([0.9, 0.5, 1.0])), area=area)
        grid = pkg.OccupancyGrid(h, w, center, res)
        grid.probs[0].copy_(torch.from_numpy(o))
        grid.image[0].copy_(torch.from_numpy(im))
        grid.update(pts, robot, **kw)
        occ.update_occupancy_map_c(o, im, pts, robot, center, res, **kw)
        assert np.array_equal(_bits(grid.probs_numpy()), _bits(o)), f"case {q}"
        assert np.array_equal(grid.image_numpy(), im), f"case {q}"
