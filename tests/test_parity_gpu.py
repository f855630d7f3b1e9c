"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star): correspondence indices bit-exact; R within 1e-5 rad and t
within 1e-5 m = 1e-2 mm of the reference's NumPy/SciPy ICP.  The kernel keeps all O(N) state
in float64, so the tests hold it to much tighter figures (stated per test).
"""
import math

import numpy as np
import pytest
import torch

from oracle import icp_oracle as orc

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5        # rad  (north star)
TRANS_TOL = 1e-2      # mm   (= 1e-5 m, north star)
TIGHT_ROT = 1e-9      # what the float64 state actually delivers
TIGHT_TRANS = 1e-6    # mm


@pytest.fixture(scope="module")
def b200():
    import icp_slam_yolo_b200 as m
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m.lib()           # fails loudly when the CUDA library is missing
    return m


def _theta(pose6):
    return np.arctan2(pose6[..., 2], pose6[..., 0])


class _Host:
    """AlignResult copied to NumPy once (one sync instead of one per pair)."""
    def __init__(self, res):
        for f in ("pose_total", "pose_last", "error", "rmse", "inliers", "iterations", "indices",
                  "src_final", "index_history"):
            v = getattr(res, f)
            setattr(self, f, None if v is None else v.cpu().numpy())


def _check_pair(res, p, o, n, *, history=True, rot=TIGHT_ROT, trans=TIGHT_TRANS):
    if not isinstance(res, _Host):
        if not hasattr(res, "_host"):
            res._host = _Host(res)
        res = res._host
    pt = res.pose_total[p]
    assert int(res.iterations[p]) == o.iterations, f"pair {p}: iterations"
    th_o = math.atan2(o.R_tot[1, 0], o.R_tot[0, 0])
    assert abs(_theta(pt) - th_o) < rot, f"pair {p}: rotation"
    assert np.max(np.abs(pt[4:6] - o.t_tot)) < trans, f"pair {p}: translation"
    assert np.allclose(pt[:4].reshape(2, 2), o.R_tot, rtol=0, atol=rot)
    if o.iterations:
        assert abs(float(res.error[p]) - o.error) < 1e-9 * max(1.0, o.error), f"pair {p}: error"
    if history and res.index_history is not None:
        h = res.index_history[p]
        for it, idx in enumerate(o.indices):
            assert np.array_equal(h[it, :n], idx), f"pair {p}: indices differ at iteration {it}"
        assert np.all(h[o.iterations:, :] == -1)
    if res.indices is not None and o.iterations:
        assert np.array_equal(res.indices[p, :n], o.indices[-1])
    if res.src_final is not None:
        assert np.allclose(res.src_final[p, :n], o.src, rtol=0, atol=trans)


# ---------------------------------------------------------------------------------------
# nearest-neighbour kernel (icp.py:37-38)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_nn_first_iteration_all_scan_pairs_bit_exact(b200, cart_scans, kernel):
    """warp-per-pair and CTA-per-pair search kernels"""
    table = b200.ScanTable.from_list(cart_scans)
    idx, d2 = b200.nn_search(table.slice_rows(1), table.slice_rows(0, table.rows - 1), kernel=kernel)
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    queries = 0
    for p in range(len(cart_scans) - 1):
        A, B = cart_scans[p + 1], cart_scans[p]
        dist, ref_idx = orc.nn_kdtree(A, B)
        assert np.array_equal(idx[p, :len(A)], ref_idx), f"pair {p}"
        assert np.all(idx[p, len(A):] == -1)
        assert np.allclose(np.sqrt(d2[p, :len(A)]), dist, rtol=1e-14, atol=0)
        queries += len(A)
    assert queries > 200000


def test_nn_near_ties_and_exact_ties(b200):
    """FP32 cannot separate these; the float64 re-decision must, lowest index on exact ties."""
    rng = np.random.default_rng(5)
    n, m = 300, 256
    B = rng.uniform(-8000, 8000, size=(m, 2))
    A = np.empty((n, 2))
    for i in range(n):          # sources almost equidistant from two targets
        a, b = rng.choice(m, 2, replace=False)
        mid = 0.5 * (B[a] + B[b])
        A[i] = mid + (B[a] - B[b]) * rng.uniform(-1e-10, 1e-10)
    B2 = np.concatenate([B, B[:64]])              # exact duplicates: ties -> lowest index
    for src, tgt in ((A, B), (A, B2), (B[:128] + 1e-7, B2)):
        s = b200.ScanTable.from_list([src])
        t = b200.ScanTable.from_list([tgt])
        idx, _ = b200.nn_search(s, t)
        _, ref = orc.nn_bruteforce(src, tgt)
        assert np.array_equal(idx[0, :len(src)].cpu().numpy(), ref)


@pytest.mark.parametrize("nn_variant", [{}, {"kernel": "warp"}, {"kernel": "warp", "dense_sweep": True},
                                        {"kernel": "cta"}])
def test_nn_shapes_ragged_and_limits(b200, nn_variant):
    rng = np.random.default_rng(11)
    cases = [(1, 1), (1, 7), (5, 1), (31, 33), (32, 8), (33, 9), (129, 257), (385, 77),
             (513, 1025), (1024, 4096), (1000, 4095)]
    for n, m in cases:
        A = rng.normal(0, 3000, size=(n, 2))
        B = rng.normal(0, 3000, size=(m, 2))
        for dt in (np.float64, np.float32):
            s = b200.ScanTable.from_list([A], dtype=dt)
            t = b200.ScanTable.from_list([B], dtype=dt)
            idx, d2 = b200.nn_search(s, t, **nn_variant)
            dist, ref = orc.nn_bruteforce(A.astype(dt), B.astype(dt))
            assert np.array_equal(idx[0].cpu().numpy(), ref), (n, m, dt)
            assert np.allclose(np.sqrt(d2[0].cpu().numpy()), dist, rtol=1e-14)


def test_nn_empty_rows_and_bad_shapes(b200):
    A = [np.zeros((0, 2)), np.array([[1.0, 2.0]]), np.array([[0.0, 0.0], [5.0, 5.0]])]
    B = [np.array([[1.0, 1.0]]), np.zeros((0, 2)), np.array([[4.0, 4.0], [1.0, 0.0]])]
    s, t = b200.ScanTable.from_list(A), b200.ScanTable.from_list(B)
    idx, d2 = b200.nn_search(s, t)
    idx = idx.cpu().numpy()
    assert np.all(idx[0] == -1) and np.all(idx[1] == -1)
    assert list(idx[2]) == [1, 0]
    big = b200.ScanTable(torch.zeros((1, 1025, 2), dtype=torch.float64, device="cuda"))
    with pytest.raises(b200.B200IcpError, match="UNSUPPORTED_SHAPE"):
        b200.nn_search(big, t.slice_rows(0, 1))
    with pytest.raises(b200.B200IcpError):
        b200.ScanTable(torch.zeros((1, 4, 2), dtype=torch.float64))       # CPU tensor: no fallback


# ---------------------------------------------------------------------------------------
# fused ICP loop (icp.py:28-53)
# ---------------------------------------------------------------------------------------
def test_circle_demo_known_answer(b200, golden):
    """icp.py:55-67.  Targets 0 and 49 of the demo coincide to 2.4e-16, so on that one
    near-exact tie SciPy's pick is undefined; poses must still agree."""
    A, B = golden["demo_A"], golden["demo_B"]
    o = orc.icp_extended(A, B)
    r = b200.icp_full(A, B)
    assert r.iterations == o.iterations == 7
    assert np.allclose(r.src, golden["demo_A_aligned"], rtol=0, atol=1e-12)
    assert np.allclose(r.R_last, golden["demo_R_est"], atol=1e-12)
    assert np.allclose(r.t_last, golden["demo_t_est"], atol=1e-12)
    same = (r.indices % 49) == (o.indices[-1] % 49)
    assert np.all(same)
    src, R, t = b200.icp(A, B)                      # reference-form return (quirk Q1)
    assert np.allclose(src, golden["demo_A_aligned"], atol=1e-12) and R.shape == (2, 2) and t.shape == (2,)


@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_scan_data_1_all_pairs_full_history(b200, cart_scans, oracle_pairs, golden, kernel):
    """Config 2: all 1,830 consecutive pairs in one launch; every iteration's correspondence
    vector, the iteration count, the pose and the error against the oracle; the pose also
    against what the unmodified reference produced (golden fixture).  Both fused kernels: the
    dispatcher picks the throughput kernel for a batch of this size, the CTA-per-pair one up to
    512 pairs."""
    table = b200.ScanTable.from_list(cart_scans)
    res = b200.align_consecutive(table, max_iterations=30, tolerance=1e-5, kernel=kernel,
                                 want_indices=True, want_src=True, want_history=True)
    torch.cuda.synchronize()
    for p, o in enumerate(oracle_pairs):
        _check_pair(res, p, o, len(cart_scans[p + 1]))
    pt = res.pose_total.cpu().numpy()
    assert np.max(np.abs(_theta(pt) - golden["pair_theta_tot"])) < TIGHT_ROT
    assert np.max(np.abs(pt[:, 4:6] - golden["pair_t_tot"])) < TIGHT_TRANS
    pl = res.pose_last.cpu().numpy()
    assert np.allclose(pl[:, :4].reshape(-1, 2, 2), golden["pair_R_last"], rtol=0, atol=1e-9)
    assert np.allclose(pl[:, 4:6], golden["pair_t_last"], rtol=0, atol=1e-6)
    # odometry chain == host prefix composition of oracle poses
    chain = b200.chain_poses(res.pose_total)
    R, t = np.eye(2), np.zeros(2)
    for o in oracle_pairs:
        t = R @ o.t_tot + t
        R = R @ o.R_tot
    assert np.allclose(chain[-1, :4].reshape(2, 2), R, atol=1e-7) and np.allclose(chain[-1, 4:], t, atol=1e-3)


@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_second_recording_all_pairs_full_history(b200, cart_scans3, oracle_pairs3, golden3, kernel):
    """scan_data_3/ (2,043 scans, 58..170 points after the filter): every consecutive pair in one
    launch; every iteration's correspondence vector, iteration count, pose and error against the
    oracle, and the pose against what the unmodified reference produced (golden fixture)."""
    table = b200.ScanTable.from_list(cart_scans3)
    res = b200.align_consecutive(table, max_iterations=30, tolerance=1e-5, kernel=kernel,
                                 want_indices=True, want_src=True, want_history=True)
    torch.cuda.synchronize()
    for p, o in enumerate(oracle_pairs3):
        _check_pair(res, p, o, len(cart_scans3[p + 1]))
    pt = res.pose_total.cpu().numpy()
    assert np.max(np.abs(_theta(pt) - golden3["pair_theta_tot"])) < TIGHT_ROT
    assert np.max(np.abs(pt[:, 4:6] - golden3["pair_t_tot"])) < TIGHT_TRANS
    pl = res.pose_last.cpu().numpy()
    assert np.allclose(pl[:, :4].reshape(-1, 2, 2), golden3["pair_R_last"], rtol=0, atol=1e-9)
    assert np.allclose(pl[:, 4:6], golden3["pair_t_last"], rtol=0, atol=1e-6)


def test_float32_inputs_synthetic_rooms(b200):
    """Config 3 shape (360 x 360, float32 tables), forced 30 iterations and tol 1e-5."""
    count = 48
    src, tgt = orc.synth_room_batch(0, count)
    s = b200.ScanTable(torch.from_numpy(src).cuda())
    t = b200.ScanTable(torch.from_numpy(tgt).cuda())
    for tol in (-1.0, 1e-5):
        res = b200.align_pairs(s, t, max_iterations=30, tolerance=tol, want_history=True, want_src=True)
        for p in range(count):
            o = orc.icp_extended(src[p], tgt[p], 30, tol)
            if tol < 0:
                assert o.iterations == 30
            _check_pair(res, p, o, 360)


def test_init_pose_and_gate(b200, cart_scans):
    pairs = [2, 99, 350, 672, 1062, 1500]
    A = [cart_scans[p + 1] for p in pairs]
    B = [cart_scans[p] for p in pairs]
    rng = np.random.default_rng(3)
    poses = []
    for _ in pairs:
        th = rng.uniform(-0.05, 0.05)
        poses.append([math.cos(th), -math.sin(th), math.sin(th), math.cos(th),
                      rng.uniform(-50, 50), rng.uniform(-50, 50)])
    init = torch.tensor(poses, dtype=torch.float64, device="cuda")
    s, t = b200.ScanTable.from_list(A), b200.ScanTable.from_list(B)
    for gate in (None, 150.0, 60.0, 1e-6):
        res = b200.align_pairs(s, t, max_iterations=30, tolerance=1e-5, init_pose=init,
                               max_corr_dist=gate, want_history=True, want_src=True)
        for k in range(len(pairs)):
            R0 = np.array(poses[k][:4]).reshape(2, 2)
            t0 = np.array(poses[k][4:])
            o = orc.icp_extended(A[k], B[k], 30, 1e-5, init_pose=(R0, t0), max_corr_dist=gate)
            _check_pair(res, k, o, len(A[k]))
            if o.iterations:
                assert abs(float(res.rmse[k]) - o.rmse) < 1e-9 * max(1.0, o.rmse)
                assert int(res.inliers[k]) == round(o.fitness * len(A[k]))
            else:
                assert math.isinf(float(res.error[k])) and int(res.inliers[k]) == 0


def test_degenerate_pairs_do_not_fail_the_batch(b200, cart_scans):
    A = [np.zeros((0, 2)), cart_scans[4], cart_scans[6], np.array([[10.0, 20.0]])]
    B = [cart_scans[3], np.zeros((0, 2)), cart_scans[5], np.array([[11.0, 22.0]])]
    s, t = b200.ScanTable.from_list(A), b200.ScanTable.from_list(B)
    res = b200.align_pairs(s, t, max_iterations=30, tolerance=1e-5, want_src=True)
    it = res.iterations.cpu().numpy()
    err = res.error.cpu().numpy()
    pt = res.pose_total.cpu().numpy()
    assert it[0] == 0 and it[1] == 0 and np.isinf(err[0]) and np.isinf(err[1])
    assert np.array_equal(pt[0], [1, 0, 0, 1, 0, 0]) and np.array_equal(pt[1], [1, 0, 0, 1, 0, 0])
    o = orc.icp_extended(A[2], B[2], 30, 1e-5)
    _check_pair(res, 2, o, len(A[2]), history=False)
    # one point onto one point: H = 0, R = I, pure translation
    o1 = orc.icp_extended(A[3], B[3], 30, 1e-5)
    assert np.allclose(pt[3, 4:6], [1.0, 2.0], atol=1e-12) and it[3] == o1.iterations == 3
    zero = b200.align_pairs(s, t, max_iterations=0, want_src=True)
    assert np.all(zero.iterations.cpu().numpy() == 0)
    assert np.allclose(zero.src_final[2, :len(A[2])].cpu().numpy(), A[2])


def test_pairings_agree(b200, cart_scans):
    rows = cart_scans[100:108]
    table = b200.ScanTable.from_list(rows)
    n = len(rows)
    tri = b200.align_pairs(table, table, pairing="triangle", max_iterations=30, tolerance=1e-5)
    count = n * (n - 1) // 2
    assert tri.pose_total.shape[0] == count
    ii, jj = [], []
    for q in range(count):
        i, j = b200.triangle_pair(q, n)
        ii.append(i); jj.append(j)
    sr = torch.tensor(jj, dtype=torch.int32, device="cuda")
    tr = torch.tensor(ii, dtype=torch.int32, device="cuda")
    exp = b200.align_pairs(table, table, pairing="explicit", src_row=sr, tgt_row=tr,
                           max_iterations=30, tolerance=1e-5)
    assert torch.equal(tri.pose_total, exp.pose_total) and torch.equal(tri.iterations, exp.iterations)
    # a shard of the triangle starting mid-way
    part = b200.align_pairs(table, table, pairing="triangle", first_pair=11, n_pairs=9,
                            max_iterations=30, tolerance=1e-5)
    assert torch.equal(part.pose_total, tri.pose_total[11:20])
    for q in (0, 5, count - 1):
        o = orc.icp_extended(rows[jj[q]], rows[ii[q]], 30, 1e-5)
        _check_pair(tri, q, o, len(rows[jj[q]]), history=False)


def test_full_size_batch_properties(b200):
    """Config 3 at full size (65,536 x 360 x 360 x 30 forced iterations): determinism, batch-
    position independence, sampled oracle parity."""
    base = 64
    src, tgt = orc.synth_room_batch(0, base)
    reps = 65536 // base
    S = torch.from_numpy(src).cuda().repeat(reps, 1, 1)
    T = torch.from_numpy(tgt).cuda().repeat(reps, 1, 1)
    s, t = b200.ScanTable(S), b200.ScanTable(T)
    r1 = b200.align_pairs(s, t, max_iterations=30, tolerance=-1.0, want_indices=True)
    r2 = b200.align_pairs(s, t, max_iterations=30, tolerance=-1.0, want_indices=True)
    torch.cuda.synchronize()
    assert torch.equal(r1.pose_total, r2.pose_total) and torch.equal(r1.indices, r2.indices)
    assert torch.all(r1.iterations == 30)
    tiled = r1.pose_total.reshape(reps, base, 6)
    assert torch.equal(tiled, tiled[0:1].expand(reps, base, 6))          # same pair, same bits
    for p in (0, 17, 63):
        o = orc.icp_extended(src[p], tgt[p], 30, -1.0)
        _check_pair(r1, 65536 - base + p, o, 360, history=False)


def test_self_alignment_is_identity(b200, cart_scans):
    table = b200.ScanTable.from_list(cart_scans[200:232])
    res = b200.align_pairs(table, table, max_iterations=30, tolerance=1e-5)
    pt = res.pose_total.cpu().numpy()
    assert np.all(res.iterations.cpu().numpy() == 1)              # quirk Q3
    assert np.allclose(pt[:, :4], [1, 0, 0, 1], atol=1e-15) and np.max(np.abs(pt[:, 4:])) < 1e-10
    assert np.all(res.error.cpu().numpy() == 0.0)


# ---------------------------------------------------------------------------------------
# scan preparation (process.py:38-52) and the reference-shaped wrappers
# ---------------------------------------------------------------------------------------
def test_polar_to_cartesian_device(b200, raw_scans, cart_scans):
    table = b200.scan_io.prepare_scans(raw_scans)
    lens = table.lengths.cpu().numpy()
    pts = table.points.cpu().numpy()
    exact = total = 0
    for k, c in enumerate(cart_scans):
        assert lens[k] == len(c), f"scan {k + 1}"
        got = pts[k, :len(c)]
        # device sin/cos are <= 2 ulp, libm's are <= 1 ulp: coordinates agree to ~1e-12 mm
        assert np.allclose(got, c, rtol=0, atol=1e-9)
        exact += int(np.sum(got == c)); total += c.size
    assert exact / total > 0.5


@pytest.mark.parametrize("variant", ["slam_offline", "realtime_2", "realtime_1"])
def test_polar_to_cartesian_reference_variants(b200, raw_scans, variant):
    """The reference's other copies of polar_to_cartesian_3d (slam_offline.py:62-75, realtime_2.py:153-165,
    realtime_1.py:160-169) differ in constants only; the device filter takes them as parameters."""
    sel = raw_scans[::37]
    table = b200.scan_io.prepare_scans(sel, filter=variant)
    lens, pts = table.lengths.cpu().numpy(), table.points.cpu().numpy()
    canon = b200.scan_io.prepare_scans(sel).lengths.cpu().numpy()
    for k, r in enumerate(sel):
        ref = orc.polar_to_cartesian_variant(r, variant)
        assert lens[k] == len(ref)
        assert np.allclose(pts[k, :len(ref)], ref[:, :2], rtol=0, atol=1e-9)
    assert int(lens.sum()) != int(canon.sum())              # the variant really filters differently


def test_reference_shaped_wrappers(b200, cart_scans):
    A, B = cart_scans[3], cart_scans[2]
    d, i = b200.nearest_neighbors(A, B)
    dr, ir = orc.nn_kdtree(A, B)
    assert np.array_equal(i, ir) and np.allclose(d, dr, rtol=1e-14)
    rmse, T = b200.registration_p2p(np.c_[A, np.zeros(len(A))], np.c_[B, np.zeros(len(B))],
                                    threshold=200.0, trans_init=np.eye(4), max_iteration=50)
    o = orc.icp_extended(A, B, 50, 1e-5, init_pose=(np.eye(2), np.zeros(2)), max_corr_dist=200.0)
    assert abs(rmse - o.rmse) < 1e-9 and np.allclose(T[:2, :2], o.R_tot, atol=1e-9)
    assert np.allclose(T[:2, 3], o.t_tot, atol=1e-6) and T.shape == (4, 4)
    r, T = b200.registration_p2p(A[:5], B)                     # gicp_lidar.py:13-15 guard
    assert math.isinf(r) and np.array_equal(T, np.eye(4))
    with pytest.raises(ValueError):
        b200.icp(np.zeros((0, 2)), B)


# ---------------------------------------------------------------------------------------
# every kernel variant, adversarial shapes for the pruned sweep
# ---------------------------------------------------------------------------------------
_VARIANTS = {
    "auto": {},
    "pair-pruned": {"kernel": "warp"},
    "pair-pruned-W1": {"kernel": "warp", "pair_warps": 1},
    "pair-pruned-W3": {"kernel": "warp", "pair_warps": 3},
    "pair-pruned-W4": {"kernel": "warp", "pair_warps": 4},
    "pair-dense": {"kernel": "warp", "dense_sweep": True},
    "pair-pruned-no-reuse": {"kernel": "warp", "sweep_reuse": False},
    "pair-dense-no-reuse": {"kernel": "warp", "dense_sweep": True, "sweep_reuse": False},
    "cta": {"kernel": "cta"},
}


@pytest.mark.parametrize("variant", list(_VARIANTS))
def test_align_variants_on_adversarial_shapes(b200, variant, cart_scans):
    """Unordered clouds (nothing to prune), far-apart scans (stage A of the pruned sweep hits
    nothing), multi-word group masks (M > 256, > 512), tiny and maximal shapes, duplicates."""
    kw = _VARIANTS[variant]
    rng = np.random.default_rng(42)
    A, B = [], []

    def add(a, b):
        A.append(np.ascontiguousarray(a, dtype=np.float64)); B.append(np.ascontiguousarray(b, dtype=np.float64))

    add(rng.normal(0, 3000, (200, 2)), rng.normal(0, 3000, (300, 2)))            # unordered
    add(rng.normal(0, 500, (90, 2)) + [40000.0, -25000.0], rng.normal(0, 500, (700, 2)))   # far apart, 3 mask words
    add(cart_scans[701], cart_scans[700] + [3000.0, 1500.0])                      # real scans, big offset
    th = np.linspace(0, 2 * np.pi, 1024, endpoint=False)
    ring = np.stack([4000 * np.cos(th), 4000 * np.sin(th)], 1)
    add(ring[::2] * 1.01 + rng.normal(0, 2, (512, 2)), np.repeat(ring, 4, axis=0)[:4096] + rng.normal(0, 1, (4096, 2)))
    add(ring + rng.normal(0, 3, ring.shape), ring[::4])                           # N = 1024 > M = 256
    add(rng.normal(0, 10, (3, 2)), rng.normal(0, 10, (2, 2)))                      # tiny
    dup = rng.normal(0, 2000, (64, 2))
    add(dup[:40] + rng.normal(0, 0.5, (40, 2)), np.concatenate([dup, dup, dup]))   # exact duplicate targets
    s, t = b200.ScanTable.from_list(A), b200.ScanTable.from_list(B)
    res = b200.align_pairs(s, t, max_iterations=12, tolerance=1e-5, want_history=True, want_src=True,
                           want_indices=True, **kw)
    if "kernel" in kw:          # the same run without per-point outputs (the LEAN instantiation for float32
        lean = b200.align_pairs(s, t, max_iterations=12, tolerance=1e-5, **kw)       # tables) gives the same bits
        assert torch.equal(lean.pose_total, res.pose_total) and torch.equal(lean.iterations, res.iterations)
        s32, t32 = b200.ScanTable.from_list(A, dtype=np.float32), b200.ScanTable.from_list(B, dtype=np.float32)
        a32 = b200.align_pairs(s32, t32, max_iterations=12, tolerance=1e-5, want_history=True, **kw)
        l32 = b200.align_pairs(s32, t32, max_iterations=12, tolerance=1e-5, **kw)
        assert torch.equal(l32.pose_total, a32.pose_total) and torch.equal(l32.error, a32.error)
        assert torch.equal(l32.iterations, a32.iterations)
    hist = res.index_history.cpu().numpy()
    for p in range(len(A)):
        o = orc.icp_extended(A[p], B[p], 12, 1e-5, nn="brute", solver="closed")
        assert np.array_equal(hist[p, 0, :len(A[p])], o.indices[0]), f"pair {p}: first search"
        if p == 1:
            # every source matches the same few targets on the near rim: H ~ 0, the rotation is
            # decided by rounding noise (parity undefined) -- only the search is checked
            continue
        _check_pair(res, p, o, len(A[p]), rot=1e-9, trans=1e-5)


def test_local_map_crop_and_dynamic_point_removal(b200, cart_scans):
    """mainn.py:300-308 and process.py:75-84 on the device, order preserved."""
    rng = np.random.default_rng(9)
    big = rng.uniform(-20000, 20000, size=(50000, 2))
    centre, radius = (1500.0, -700.0), 6000.0
    got = b200.crop_local_map(torch.from_numpy(big).cuda(), centre, radius).cpu().numpy()
    ref = big[np.sum((big - np.array(centre)) ** 2, axis=1) < radius ** 2]
    assert np.array_equal(got, ref)
    tiny = b200.crop_local_map(torch.from_numpy(big).cuda(), (1e6, 1e6), 10.0)
    assert tiny.shape[0] == len(big)                               # < 50 survivors: whole map
    cur, prev = cart_scans[500], cart_scans[499]
    moved = cur.copy(); moved[::7] += 400.0                        # "dynamic" points
    d, _ = orc.nn_kdtree(moved, prev)
    for dt in (np.float64, np.float32):
        got = b200.remove_dynamic_points(torch.from_numpy(moved.astype(dt)).cuda(),
                                         torch.from_numpy(prev.astype(dt)).cuda(), 250.0).cpu().numpy()
        dd, _ = orc.nn_kdtree(moved.astype(dt).astype(np.float64), prev.astype(dt).astype(np.float64))
        assert np.array_equal(got, moved.astype(dt)[dd < 250.0])
    assert 0 < np.sum(d < 250.0) < len(moved)
    # large sets go through the sharded-map search
    mp = orc.synth_map(30000, dtype=np.float64); sc = orc.synth_scan_for_map(2000, dtype=np.float64)
    got = b200.remove_dynamic_points(torch.from_numpy(sc).cuda(), torch.from_numpy(mp).cuda(), 30.0).cpu().numpy()
    dd, _ = orc.nn_kdtree(sc, mp)
    assert np.array_equal(got, sc[dd < 30.0])
    assert b200.remove_dynamic_points(torch.from_numpy(sc).cuda(), None).shape[0] == len(sc)


def test_voxel_down_sample_and_gicp_shaped_call(b200, cart_scans):
    """gicp_lidar.py:8-11,20-21 (voxel_down_sample before registration), d.py:10-16 (2D grid)."""
    rng = np.random.default_rng(2)
    for pts, vox in ((cart_scans[800], 50.0), (rng.uniform(-3000, 3000, (20000, 2)), 70.0),
                     (np.array([[0.1, 0.1], [0.2, 0.3], [-0.1, 0.1], [5.0, 5.0]]), 1.0)):
        ref = orc.voxel_down_sample_2d(pts, vox)
        for dt, tol in ((np.float64, 1e-9), (np.float32, 1e-3)):
            got = b200.voxel_down_sample(torch.from_numpy(pts.astype(dt)).cuda(), vox).cpu().numpy()
            ref_dt = orc.voxel_down_sample_2d(pts.astype(dt), vox)
            assert got.shape == ref_dt.shape and np.allclose(got, ref_dt, rtol=0, atol=tol)
        assert len(ref) <= len(pts)
    A, B = cart_scans[801], cart_scans[800]
    rmse, T = b200.registration_p2p(np.c_[A, np.zeros(len(A))], np.c_[B, np.zeros(len(B))], 200.0, 20.0, np.eye(4))
    a, b_ = orc.voxel_down_sample_2d(A, 20.0), orc.voxel_down_sample_2d(B, 20.0)
    o = orc.icp_extended(a, b_, 50, 1e-5, init_pose=(np.eye(2), np.zeros(2)), max_corr_dist=200.0)
    assert abs(rmse - o.rmse) < 1e-7 and np.allclose(T[:2, :2], o.R_tot, atol=1e-9) and np.allclose(T[:2, 3], o.t_tot, atol=1e-5)


def test_host_pipeline_matches_resident_path(b200, cart_scans):
    """HostPipeline (pinned host tables, chunked H2D || kernel || D2H) == align_pairs on resident
    tensors, ragged rows included, chunk boundaries not multiples of anything."""
    A = [cart_scans[p + 1] for p in range(300, 437)]
    B = [cart_scans[p] for p in range(300, 437)]
    hs, hsl = b200.ScanTable.pack_host(A, dtype=np.float64, pin=True)
    ht, htl = b200.ScanTable.pack_host(B, dtype=np.float64, pin=True)
    pipe = b200.registration.HostPipeline(len(A), hs.shape[1], ht.shape[1], dtype=torch.float64, chunks=5)
    pose, err, its = pipe.run(hs, ht, hsl, htl, max_iterations=30, tolerance=1e-5)
    torch.cuda.synchronize()
    ref = b200.align_pairs(b200.ScanTable(hs.cuda(), hsl.cuda()), b200.ScanTable(ht.cuda(), htl.cuda()),
                           max_iterations=30, tolerance=1e-5)
    assert torch.equal(pose, ref.pose_total.cpu()) and torch.equal(err, ref.error.cpu())
    assert torch.equal(its, ref.iterations.cpu()) and pipe.launches == len(pipe.bounds) - 1
    pose2, _, _ = pipe.run(hs, ht, hsl, htl, max_iterations=30, tolerance=1e-5)      # buffers are reusable
    torch.cuda.synchronize()
    assert torch.equal(pose2, ref.pose_total.cpu())


def test_host_pipeline_back_to_back_runs_do_not_race(b200):
    """Two run() calls with DIFFERENT inputs and no synchronisation in between: the second run's
    first copies must not land in staging buffers the first run's kernels are still reading."""
    n = 6000
    s1, t1 = orc.synth_room_batch(100, n)
    s2, t2 = orc.synth_room_batch(50000, n)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h = [pin(s1), pin(t1), pin(s2), pin(t2)]
    pipe = b200.registration.HostPipeline(n, 360, 360, dtype=torch.float32, chunks=4, graph=True)
    for rep in range(3):
        pa, _, _ = pipe.run(h[0], h[1], max_iterations=30, tolerance=-1.0)
        if rep == 0:                                        # keep the first run's result once (needs a sync)
            pipe.done_event.synchronize()
            first = pa.clone()
        pb, _, ib = pipe.run(h[2], h[3], max_iterations=30, tolerance=-1.0)
        pipe.done_event.synchronize()
        second = pb.clone()
        ref2 = b200.align_pairs(b200.ScanTable(h[2].cuda()), b200.ScanTable(h[3].cuda()), max_iterations=30,
                                tolerance=-1.0, kernel="warp")
        assert torch.equal(second, ref2.pose_total.cpu()), f"repetition {rep}: second run corrupted"
    ref1 = b200.align_pairs(b200.ScanTable(h[0].cuda()), b200.ScanTable(h[1].cuda()), max_iterations=30,
                            tolerance=-1.0, kernel="warp")
    assert torch.equal(first, ref1.pose_total.cpu())
    assert len(pipe._graphs) == 2             # repetitions 1 and 2 were CUDA-graph captures / replays of both runs
    pa, _, _ = pipe.run(h[0], h[1], max_iterations=30, tolerance=-1.0)                  # a replay, checked on its own
    pipe.done_event.synchronize()
    assert torch.equal(pa, ref1.pose_total.cpu())
    eager = b200.registration.HostPipeline(n, 360, 360, dtype=torch.float32, chunks=4, graph=False)
    for _ in range(2):
        pe, _, _ = eager.run(h[0], h[1], max_iterations=30, tolerance=-1.0)
    eager.done_event.synchronize()
    assert torch.equal(pe, ref1.pose_total.cpu()) and not eager._graphs
    with pytest.raises(ValueError):
        pipe.run(h[0], h[1], torch.zeros(n, dtype=torch.int32).pin_memory(), None)      # one length array only
    with pytest.raises(ValueError):
        pipe.run(torch.from_numpy(s1), h[1])                                             # not pinned


@pytest.mark.parametrize("tol", [-1.0, 1e-5])
def test_benchmark_batch_sample_full_history_vs_oracle(b200, tol):
    """512 pairs drawn from the actual 65,536-pair benchmark batch (bench.py: synth_room_batch(0, 65536)):
    every iteration's correspondence vector, the iteration count, pose and error against the oracle,
    through the throughput kernel, for the forced-30 and the tolerance-1e-5 runs."""
    rng = np.random.default_rng(7)
    picks = np.sort(rng.choice(65536, size=512, replace=False))
    src = np.concatenate([orc.synth_room_batch(int(q), 1)[0] for q in picks])
    tgt = np.concatenate([orc.synth_room_batch(int(q), 1)[1] for q in picks])
    s, t = b200.ScanTable(torch.from_numpy(src).cuda()), b200.ScanTable(torch.from_numpy(tgt).cuda())
    res = b200.align_pairs(s, t, max_iterations=30, tolerance=tol, kernel="warp", want_history=True, want_indices=True)
    lean = b200.align_pairs(s, t, max_iterations=30, tolerance=tol, kernel="warp")
    assert torch.equal(lean.pose_total, res.pose_total) and torch.equal(lean.iterations, res.iterations)
    for p in range(len(picks)):
        o = orc.icp_extended(src[p], tgt[p], 30, tol)
        _check_pair(res, p, o, 360)


def test_triangle_pairing_trajectory_scans_vs_oracle(b200):
    """configs[3] shape: synth_trajectory_scans (256 rows), pairing='triangle' with a mid-range
    first_pair (a shard that starts inside a row of the triangle), against the oracle pair by pair."""
    rows = 256
    scans = orc.synth_trajectory_scans(rows)
    table = b200.ScanTable(torch.from_numpy(scans).cuda())
    first, count = 12345, 300
    res = b200.align_pairs(table, table, pairing="triangle", first_pair=first, n_pairs=count, max_iterations=30,
                           tolerance=1e-5, kernel="warp", want_history=True, want_indices=True)
    for q in range(count):
        i, j = b200.triangle_pair(first + q, rows)
        assert 0 <= i < j < rows
        o = orc.icp_extended(scans[j], scans[i], 30, 1e-5)
        _check_pair(res, q, o, 360)
    last = b200.triangle_pair_count(rows) - 5            # the tail of the enumeration
    tail = b200.align_pairs(table, table, pairing="triangle", first_pair=last, n_pairs=5, max_iterations=30,
                            tolerance=-1.0, kernel="warp", want_history=True)
    for q in range(5):
        i, j = b200.triangle_pair(last + q, rows)
        _check_pair(tail, q, orc.icp_extended(scans[j], scans[i], 30, -1.0), 360)


def test_dense_sweep_flag_gives_identical_bits(b200, cart_scans):
    table = b200.ScanTable.from_list(cart_scans[900:964])
    a = b200.align_consecutive(table, max_iterations=30, tolerance=1e-5, want_indices=True, want_stats=True)
    d = b200.align_consecutive(table, max_iterations=30, tolerance=1e-5, want_indices=True, want_stats=True,
                               dense_sweep=True)
    assert torch.equal(a.pose_total, d.pose_total) and torch.equal(a.indices, d.indices)
    assert torch.equal(a.iterations, d.iterations) and torch.equal(a.error, d.error)
    assert int(a.evaluated_pairs.sum()) < int(d.evaluated_pairs.sum())      # something was culled


def test_sweep_reuse_gives_identical_bits_and_skips_sweeps(b200, cart_scans):
    """Passes skip their candidate sweep while the points provably keep their group: same
    correspondences at every iteration, same poses bit for bit, far fewer evaluated pairs --
    on real scans (early exit) and on synthetic rooms with 30 forced iterations."""
    table = b200.ScanTable.from_list(cart_scans[900:1028])
    kw = dict(max_iterations=30, tolerance=1e-5, want_indices=True, want_stats=True, want_history=True)
    a = b200.align_consecutive(table, **kw)
    d = b200.align_consecutive(table, sweep_reuse=False, **kw)
    assert torch.equal(a.index_history, d.index_history) and torch.equal(a.pose_total, d.pose_total)
    assert torch.equal(a.iterations, d.iterations) and torch.equal(a.error, d.error)
    assert int(a.evaluated_pairs.sum()) < int(d.evaluated_pairs.sum())
    src, tgt = orc.synth_room_batch(4000, 64)
    s, t = b200.ScanTable(torch.from_numpy(src).cuda()), b200.ScanTable(torch.from_numpy(tgt).cuda())
    kw = dict(max_iterations=30, tolerance=-1.0, want_stats=True, want_history=True)
    a = b200.align_pairs(s, t, **kw)
    d = b200.align_pairs(s, t, sweep_reuse=False, **kw)
    assert torch.equal(a.index_history, d.index_history) and torch.equal(a.pose_total, d.pose_total)
    assert torch.equal(a.error, d.error)
    assert int(a.evaluated_pairs.sum()) * 2 < int(d.evaluated_pairs.sum())       # most sweeps are skipped


def test_best_fit_transform_matches_reference_form(b200, cart_scans, golden):
    """icp.py:5-26 on matched rows; the demo pair and real scans; batched ragged form."""
    R, t = b200.best_fit_transform(golden["demo_A"], golden["demo_A_aligned"])
    Ro, to = orc.best_fit_transform(golden["demo_A"], golden["demo_A_aligned"])
    assert np.allclose(R, Ro, atol=1e-12) and np.allclose(t, to, atol=1e-12)
    P = [cart_scans[k][:100] for k in (5, 40, 300)] + [cart_scans[7][:3]]
    Q = [cart_scans[k + 1][:100] for k in (5, 40, 300)] + [cart_scans[8][:3]]
    poses = b200.best_fit(b200.ScanTable.from_list(P), b200.ScanTable.from_list(Q)).cpu().numpy()
    for k in range(len(P)):
        Ro, to = orc.best_fit_transform(P[k], Q[k])
        assert np.allclose(poses[k, :4].reshape(2, 2), Ro, atol=1e-11) and np.allclose(poses[k, 4:], to, atol=1e-7)
    with pytest.raises(ValueError):
        b200.best_fit_transform(P[0], Q[0][:50])


def test_icp_drop_in_handles_large_sets(b200):
    """A scan against a local map larger than the fused kernel's tile (the reference's saved map has
    11,283 points, global_map_offline.pcd): icp() routes through the sharded-map path."""
    mp = orc.synth_map(11283, dtype=np.float64)
    sc = orc.synth_scan_for_map(1500, dtype=np.float64)
    o = orc.icp_extended(sc, mp, 30, 1e-5)
    r = b200.icp_full(sc, mp, 30, 1e-5)
    assert r.iterations == o.iterations and np.array_equal(r.indices, o.indices[-1])
    assert np.allclose(r.R, o.R_tot, atol=1e-9) and np.allclose(r.t, o.t_tot, atol=1e-6)
    assert np.allclose(r.src, o.src, atol=1e-6) and abs(r.error - o.error) < 1e-9 * max(1.0, o.error)
    src, R, t = b200.icp(sc, mp, 30, 1e-5)
    assert np.allclose(R, o.R_last, atol=1e-9) and np.allclose(t, o.t_last, atol=1e-6)
    g = b200.icp_full(sc, mp, 30, 1e-5, init_pose=(np.eye(2), np.array([3.0, -2.0])), max_corr_dist=60.0)
    og = orc.icp_extended(sc, mp, 30, 1e-5, init_pose=(np.eye(2), np.array([3.0, -2.0])), max_corr_dist=60.0)
    assert g.iterations == og.iterations and abs(g.rmse - og.rmse) < 1e-9 * max(1.0, og.rmse)
    assert abs(g.fitness - og.fitness) < 1e-12


@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_nn_randomised_stress_exact_indices(b200, kernel):
    """2,400 random ragged problems: clustered, lattice (exact ties -> lowest index), collinear,
    duplicated, huge offsets (1e6) and tiny scales (1e-3): the pruned FP32 sweep + float64
    re-decision must equal the float64 brute-force argmin everywhere."""
    rng = np.random.default_rng(2024)
    A, B = [], []
    for q in range(2400):
        n, m = int(rng.integers(1, 200)), int(rng.integers(1, 260))
        kind = q % 6
        scale = [1.0, 1e3, 1e-3, 50.0, 1e4, 7.0][q % 6 if q % 12 < 6 else (q + 1) % 6]
        off = rng.uniform(-1, 1, 2) * (1e6 if q % 7 == 0 else 1e3)
        if kind == 0:                       # gaussian blobs
            a, b = rng.normal(0, 1, (n, 2)), rng.normal(0, 1, (m, 2))
        elif kind == 1:                     # integer lattice: many exact ties
            a, b = rng.integers(-6, 7, (n, 2)).astype(float), rng.integers(-6, 7, (m, 2)).astype(float)
        elif kind == 2:                     # collinear
            a = np.stack([np.sort(rng.uniform(-5, 5, n)), np.zeros(n)], 1)
            b = np.stack([np.sort(rng.uniform(-5, 5, m)), rng.normal(0, 1e-9, m)], 1)
        elif kind == 3:                     # duplicates of a few sites
            sites = rng.normal(0, 3, (5, 2))
            a, b = sites[rng.integers(0, 5, n)] + rng.normal(0, 0.01, (n, 2)), sites[rng.integers(0, 5, m)]
        elif kind == 4:                     # ordered arcs (prunable)
            ta, tb = np.sort(rng.uniform(0, 6.28, n)), np.sort(rng.uniform(0, 6.28, m))
            a = np.stack([np.cos(ta), np.sin(ta)], 1) * 3 + rng.normal(0, 0.01, (n, 2))
            b = np.stack([np.cos(tb), np.sin(tb)], 1) * 3
        else:                               # far apart clusters
            a, b = rng.normal(0, 0.3, (n, 2)) + 40.0, rng.normal(0, 0.3, (m, 2))
        A.append(a * scale + off); B.append(b * scale + off)
    for dt in (np.float64, np.float32):
        s, t = b200.ScanTable.from_list(A, dtype=dt), b200.ScanTable.from_list(B, dtype=dt)
        idx, d2 = b200.nn_search(s, t, kernel=kernel)
        idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
        hs, ht = s.points.cpu().numpy().astype(np.float64), t.points.cpu().numpy().astype(np.float64)
        for q in range(len(A)):
            n, m = len(A[q]), len(B[q])
            dd = ((hs[q, :n, None, :] - ht[q, None, :m, :]) ** 2).sum(axis=2)
            ref = dd.argmin(axis=1)
            assert np.array_equal(idx[q, :n], ref), f"problem {q} ({dt.__name__}, kind {q % 6})"
            assert np.array_equal(d2[q, :n], dd[np.arange(n), ref])
