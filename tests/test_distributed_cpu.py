"""CPU, world_size 2, gloo: the multi-rank host logic (no GPU, no kernels).

* pair workloads: contiguous shard ranges cover every pair exactly once, the per-rank results
  gathered in rank order equal the single-process result (the oracle stands in for the kernel);
* scan-to-map protocol: per-rank records (distance, GLOBAL index, matched point) all-gathered and
  merged with "smaller distance, then lower global index" reproduce the global nearest neighbour.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import icp_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import icp_slam_yolo_b200 as m
    # ---- pair sharding
    n_pairs = 13
    b, e = m.shard_range(n_pairs, rank, world)
    poses = np.zeros((e - b, 6))
    for k, p in enumerate(range(b, e)):
        s, t, _, _ = orc.synth_room_pair(p, n_points=90)
        r = orc.icp_extended(s, t, 10, 1e-5, keep_history=False)
        poses[k] = np.concatenate([r.R_tot.reshape(4), r.t_tot])
    sizes = [m.shard_range(n_pairs, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)               # ranges differ by at most one pair
    mine = torch.zeros((cap, 6), dtype=torch.float64)
    mine[: e - b] = torch.from_numpy(poses)
    gathered = [torch.zeros((cap, 6), dtype=torch.float64) for _ in sizes]
    dist.all_gather(gathered, mine)
    allposes = torch.cat([g[: hi - lo] for g, (lo, hi) in zip(gathered, sizes)]).numpy()
    # ---- timing reduction used by bench.py: max over ranks
    tmax = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    # ---- scan-to-map record protocol
    map_pts = orc.synth_map(6000, dtype=np.float64)
    map_pts = np.concatenate([map_pts, map_pts[:500]])          # duplicates across shards: ties
    scan = orc.synth_scan_for_map(300, dtype=np.float64)
    lo, hi = m.shard_range(len(map_pts), rank, world)
    dloc, iloc = orc.nn_bruteforce(scan, map_pts[lo:hi])
    rec = np.stack([dloc ** 2, (iloc + lo).astype(np.float64), map_pts[lo:hi][iloc, 0], map_pts[lo:hi][iloc, 1]], axis=1)
    # round-2 protocol: a rank answers "none" (+inf) for every point whose nearest neighbour provably
    # lies in another shard, i.e. whose local best is farther than a bound of the GLOBAL distance
    # (device: bounding circles of all ranks' chunks; here: the min over ranks of the local bests)
    ub = torch.from_numpy(dloc.copy())
    dist.all_reduce(ub, op=dist.ReduceOp.MIN)
    none = dloc > ub.numpy()
    rec[none] = [np.inf, float(np.iinfo(np.int64).max), 0.0, 0.0]
    rec_all = torch.zeros((world * len(scan), 4), dtype=torch.float64)    # rank-major, like NCCL
    dist.all_gather_into_tensor(rec_all, torch.from_numpy(rec))
    ra = rec_all.numpy().reshape(world, len(scan), 4)
    order = np.lexsort((ra[:, :, 1], ra[:, :, 0]), axis=0)[0]   # distance, then global index
    win = ra[order, np.arange(len(scan))]
    if rank == 0:
        np.savez(os.path.join(out_dir, "r0.npz"), poses=allposes, tmax=tmax.numpy(), win=win)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "r0.npz")
    assert z["tmax"][0] == 11.0
    ref = []
    for p in range(13):
        s, t, _, _ = orc.synth_room_pair(p, n_points=90)
        r = orc.icp_extended(s, t, 10, 1e-5, keep_history=False)
        ref.append(np.concatenate([r.R_tot.reshape(4), r.t_tot]))
    assert np.array_equal(z["poses"], np.array(ref))
    map_pts = orc.synth_map(6000, dtype=np.float64)
    map_pts = np.concatenate([map_pts, map_pts[:500]])
    scan = orc.synth_scan_for_map(300, dtype=np.float64)
    d, i = orc.nn_bruteforce(scan, map_pts)
    assert np.array_equal(z["win"][:, 1].astype(np.int64), i)
    assert np.allclose(z["win"][:, 0], d ** 2, rtol=1e-15) and np.array_equal(z["win"][:, 2:], map_pts[i])
