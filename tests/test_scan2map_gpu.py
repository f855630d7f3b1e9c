"""GPU parity of the scan-to-map path (config 5 at reduced size) against the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import icp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import icp_slam_yolo_b200 as m
    assert torch.cuda.is_available()
    m.lib()
    return m


def _shards(b200, map_pts, n_shards):
    out = []
    for g in range(n_shards):
        b, e = b200.shard_range(len(map_pts), g, n_shards)
        out.append(b200.MapShard(torch.from_numpy(map_pts[b:e]).cuda(), global_offset=b))
    return out


@pytest.mark.parametrize("m_points,n_scan,n_shards,dtype", [
    (1 << 18, 8192, 4, np.float32),      # parity on the reduced instance M = 2^18
    (50000, 1000, 3, np.float64),        # ragged: shard sizes not multiples of the chunk
    (777, 130, 1, np.float64),           # a single partial chunk
])
def test_scan_to_map_matches_oracle_every_iteration(b200, m_points, n_scan, n_shards, dtype):
    map_pts = orc.synth_map(m_points, dtype=dtype)
    scan = orc.synth_scan_for_map(n_scan, dtype=dtype)
    iters = 6
    o = orc.icp_extended(scan, map_pts, iters, -1.0)            # KD-tree oracle, forced iterations
    run = b200.scan_to_map.ScanToMapLocalShards(_shards(b200, map_pts, n_shards), n_scan)
    run.init(torch.from_numpy(scan).cuda())
    for it in range(iters):
        run.step(iters, -1.0)
        idx = run.indices.cpu().numpy()
        assert np.array_equal(idx, o.indices[it]), f"iteration {it}: {np.sum(idx != o.indices[it])} indices differ"
    r = run.result()
    assert r.iterations == iters
    assert abs(math.atan2(r.R[1, 0], r.R[0, 0]) - math.atan2(o.R_tot[1, 0], o.R_tot[0, 0])) < 1e-9
    assert np.max(np.abs(r.t - o.t_tot)) < 1e-6 and abs(r.error - o.error) < 1e-9 * max(1.0, o.error)
    assert np.allclose(r.src.cpu().numpy(), o.src, rtol=0, atol=1e-6)


def test_scan_to_map_convergence_gate_and_init_pose(b200):
    map_pts = orc.synth_map(60000, dtype=np.float64)
    scan = orc.synth_scan_for_map(2000, dtype=np.float64)
    th = 0.01
    init = [math.cos(th), -math.sin(th), math.sin(th), math.cos(th), 10.0, -5.0]
    for gate in (None, 40.0):
        o = orc.icp_extended(scan, map_pts, 25, 1e-5, init_pose=(np.array(init[:4]).reshape(2, 2), np.array(init[4:])),
                             max_corr_dist=gate)
        shard = b200.MapShard(torch.from_numpy(map_pts).cuda())
        r = b200.scan_to_map_icp(torch.from_numpy(scan).cuda(), shard, 25, 1e-5, init_pose=init,
                                 max_corr_dist=gate, want_indices=True)
        assert r.iterations == o.iterations and r.inliers == round(o.fitness * len(scan))
        assert np.array_equal(r.indices.cpu().numpy(), o.indices[-1])
        assert np.allclose(r.R, o.R_tot, atol=1e-9) and np.allclose(r.t, o.t_tot, atol=1e-6)
        assert abs(r.error - o.error) < 1e-9 * max(1.0, o.error) and abs(r.rmse - o.rmse) < 1e-9 * max(1.0, o.rmse)
        assert np.allclose(r.R_last, o.R_last, atol=1e-9) and np.allclose(r.t_last, o.t_last, atol=1e-6)


def test_scan_to_map_duplicates_resolve_to_lowest_global_index(b200):
    """Exact float64 ties across shards: the winner is the lowest GLOBAL index."""
    rng = np.random.default_rng(1)
    base = rng.uniform(-5000, 5000, size=(3000, 2))
    map_pts = np.concatenate([base, base, base])              # every point exists in 3 shards
    scan = base[:512] + rng.normal(0, 1.0, size=(512, 2))
    run = b200.scan_to_map.ScanToMapLocalShards(_shards(b200, map_pts, 3), 512)
    run.init(torch.from_numpy(scan).cuda())
    run.step(1, -1.0)
    _, ref = orc.nn_bruteforce(scan, map_pts)
    assert np.array_equal(run.indices.cpu().numpy(), ref) and ref.max() < 3000


def test_scan_to_map_cuda_graph_replay_equals_eager(b200):
    """ScanToMap(graph=True): the whole fixed-length loop captured once and replayed (different
    scans, same arguments) gives the bits of the eager loop."""
    map_pts = orc.synth_map(1 << 16)
    shard = b200.MapShard(torch.from_numpy(map_pts).cuda())
    eager = b200.ScanToMap(shard, 2048, want_indices=True)
    graph = b200.ScanToMap(shard, 2048, want_indices=True, graph=True)
    for seed in (77, 78, 79):
        scan = torch.from_numpy(orc.synth_scan_for_map(2048, scan_seed=seed)).cuda()
        for tol in (-1.0, 1e-2):
            a = eager.run(scan, max_iterations=12, tolerance=tol)
            ia, sa = a.indices.clone(), a.src.clone()
            g = graph.run(scan, max_iterations=12, tolerance=tol)
            assert a.iterations == g.iterations and a.error == g.error
            assert np.array_equal(a.R, g.R) and np.array_equal(a.t, g.t)
            assert torch.equal(ia, g.indices) and torch.equal(sa, g.src)
    o = orc.icp_extended(orc.synth_scan_for_map(2048, scan_seed=79), map_pts, 12, 1e-2)
    assert g.iterations == o.iterations and np.array_equal(g.indices.cpu().numpy(), o.indices[-1])


def test_scan_to_map_float64_map_and_unaligned_sizes(b200):
    """float64 map (16-byte points: the four-load scan path), shard sizes that leave a partial last
    chunk and a padded circle table, scan sizes that are not a multiple of the CTA's eight points."""
    map_pts = orc.synth_map(40000 + 17, dtype=np.float64)
    scan = orc.synth_scan_for_map(1003, dtype=np.float64)
    o = orc.icp_extended(scan, map_pts, 5, -1.0)
    for n_shards in (1, 5):
        run = b200.scan_to_map.ScanToMapLocalShards(_shards(b200, map_pts, n_shards), len(scan))
        run.init(torch.from_numpy(scan).cuda())
        for it in range(5):
            run.step(5, -1.0)
            assert np.array_equal(run.indices.cpu().numpy(), o.indices[it]), (n_shards, it)
        r = run.result()
        assert np.allclose(r.R, o.R_tot, atol=1e-9) and np.allclose(r.t, o.t_tot, atol=1e-6)


def test_scan_to_map_spatial_sort_makes_unordered_maps_cullable(b200):
    """A map in arbitrary order (what a hash-based voxel filter leaves behind): "auto" builds a
    Morton-sorted copy per shard; indices still refer to the ORIGINAL order and exact ties (duplicated
    points) still resolve to the lowest original global index, at every iteration."""
    rng = np.random.default_rng(12)
    ordered = orc.synth_map(60000, dtype=np.float64)
    assert b200.MapShard(torch.from_numpy(ordered).cuda()).order is None            # already coherent: scanned as given
    shuffled = ordered[rng.permutation(len(ordered))]
    map_pts = np.concatenate([shuffled, shuffled[:700]])                            # exact duplicates, far apart in memory
    scan = orc.synth_scan_for_map(1000, dtype=np.float64)
    o = orc.icp_extended(scan, map_pts, 5, -1.0, nn="brute", solver="closed")       # lowest index on ties
    for n_shards, dtype in ((1, np.float64), (3, np.float64), (2, np.float32)):
        mp = map_pts.astype(dtype)
        oo = o if dtype == np.float64 else orc.icp_extended(scan.astype(dtype), mp, 5, -1.0, nn="brute", solver="closed")
        shards = _shards(b200, mp, n_shards)
        assert all(s.order is not None for s in shards)
        run = b200.scan_to_map.ScanToMapLocalShards(shards, len(scan))
        run.init(torch.from_numpy(scan.astype(dtype)).cuda())
        for it in range(5):
            run.step(5, -1.0)
            assert np.array_equal(run.indices.cpu().numpy(), oo.indices[it]), (n_shards, dtype, it)
        r = run.result()
        assert np.allclose(r.R, oo.R_tot, atol=1e-9) and np.allclose(r.t, oo.t_tot, atol=1e-6)
    # forced on an ordered map: same answers as the unsorted scan
    a = b200.scan_to_map_icp(torch.from_numpy(scan).cuda(), b200.MapShard(torch.from_numpy(ordered).cuda(), spatial_sort=True),
                             6, -1.0, want_indices=True)
    b = b200.scan_to_map_icp(torch.from_numpy(scan).cuda(), b200.MapShard(torch.from_numpy(ordered).cuda(), spatial_sort=False),
                             6, -1.0, want_indices=True)
    assert torch.equal(a.indices, b.indices) and a.error == b.error and np.array_equal(a.R, b.R)


def test_scan_to_map_ties_in_given_order_closest_chunk_first(b200):
    """Exact duplicates INSIDE one shard, in different chunks of a map scanned in its given order.
    The search scans the chunk with the closest centre first and the others in the order the CTA's
    super-circle pass left them, so the tie rule (lowest original index) must not depend on the
    scanning order; and from the second iteration on the FP32 filter starts from the distance to the
    previous nearest neighbour, which every duplicate of that point must pass."""
    wall = orc.synth_map(24000, dtype=np.float32)
    a, b = wall[:9000], wall[9000:24000]
    # the copy of `a` starts at index 24000 - not a multiple of the 1,024-point chunk relative to `a`
    # (9000 + 15000 = 24000 = 23 chunks + 448): its chunks, circles and centres differ from the original's
    map_pts = np.concatenate([a, b, a]).astype(np.float32)
    scan = orc.synth_scan_for_map(1500, dtype=np.float32)
    o = orc.icp_extended(scan, map_pts, 6, -1.0, nn="brute", solver="closed")       # lowest index on ties
    assert (o.indices[-1] < 9000).any()                                             # ties are exercised
    for n_shards in (1, 2):
        shards = []
        for g in range(n_shards):
            lo, hi = b200.shard_range(len(map_pts), g, n_shards)
            shards.append(b200.MapShard(torch.from_numpy(map_pts[lo:hi]).cuda(), global_offset=lo,
                                        spatial_sort=False))                        # scanned as given
        run = b200.scan_to_map.ScanToMapLocalShards(shards, len(scan))
        run.init(torch.from_numpy(scan).cuda())
        for it in range(6):
            run.step(6, -1.0)
            idx = run.indices.cpu().numpy()
            assert np.array_equal(idx, o.indices[it]), (n_shards, it, int(np.sum(idx != o.indices[it])))
            assert idx.max() < 24000                                                # never the later copy
