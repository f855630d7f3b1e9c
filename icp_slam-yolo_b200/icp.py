"""Drop-in mirror of the reference's ICP call surface, executed on the B200.

Same names, argument meaning and return shapes as
``labels_segmentation/icp.py`` (``icp`` :28-53, ``best_fit_transform`` :5-26) and the
Open3D-shaped wrapper ``gicp`` (duc/ICP_LIDAR/gicp_lidar.py:12-36), with NumPy arrays in
and out like the reference.  All arithmetic runs in the CUDA library through the C ABI;
nothing here computes on the host and nothing falls back to the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _cabi
from .registration import ScanTable, align_pairs, best_fit, nn_search


@dataclass
class IcpOutput:
    R: np.ndarray            # (2,2) cumulative rotation,  src = R A + t
    t: np.ndarray            # (2,)
    error: float             # mean NN distance of the last search (lagged, icp.py:48)
    iterations: int
    R_last: np.ndarray       # last increment (what the reference's icp() returns)
    t_last: np.ndarray
    rmse: float
    fitness: float
    indices: np.ndarray      # (N,) correspondences of the last search
    src: np.ndarray          # (N,2) transformed source


def _as_points(a, name: str) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] < 2:
        raise ValueError(f"{name} must be an (N, 2) array, got shape {a.shape}")
    if a.shape[0] == 0:
        # the reference fails inside SciPy on empty input
        raise ValueError(f"{name} is empty")
    return np.ascontiguousarray(a[:, :2])


def _pose6(init_pose) -> Optional[torch.Tensor]:
    if init_pose is None:
        return None
    if isinstance(init_pose, (tuple, list)) and len(init_pose) == 2:
        R0, t0 = np.asarray(init_pose[0], dtype=np.float64), np.asarray(init_pose[1], dtype=np.float64)
    else:
        T = np.asarray(init_pose, dtype=np.float64)
        if T.shape == (4, 4):            # Open3D trans_init (gicp_lidar.py:12)
            R0, t0 = T[:2, :2], T[:2, 3]
        elif T.shape == (3, 3):
            R0, t0 = T[:2, :2], T[:2, 2]
        else:
            raise ValueError("init_pose must be (R, t), a 3x3 or a 4x4 homogeneous matrix")
    p = np.concatenate([R0.reshape(4), t0.reshape(2)])[None, :]
    return torch.from_numpy(np.ascontiguousarray(p)).cuda()


def icp_full(A, B, max_iterations: int = 20, tolerance: float = 1e-5, *, init_pose=None,
             max_corr_dist: Optional[float] = None) -> IcpOutput:
    """One alignment with every output of the new call surface."""
    A = _as_points(A, "A")
    B = _as_points(B, "B")
    lib = _cabi.lib()
    if len(A) > lib.b200icp_max_src_pitch() or len(B) > lib.b200icp_max_tgt_pitch():
        return _icp_full_large(A, B, max_iterations, tolerance, init_pose, max_corr_dist)
    src = ScanTable(torch.from_numpy(A[None]).cuda(), None)
    tgt = ScanTable(torch.from_numpy(B[None]).cuda(), None)
    res = align_pairs(src, tgt, n_pairs=1, max_iterations=max_iterations, tolerance=tolerance,
                      init_pose=_pose6(init_pose), max_corr_dist=max_corr_dist,
                      want_indices=True, want_src=True)
    pt = res.pose_total[0].cpu().numpy()
    pl = res.pose_last[0].cpu().numpy()
    inl = int(res.inliers[0].item())
    return IcpOutput(
        R=pt[:4].reshape(2, 2).copy(), t=pt[4:6].copy(),
        error=float(res.error[0].item()), iterations=int(res.iterations[0].item()),
        R_last=pl[:4].reshape(2, 2).copy(), t_last=pl[4:6].copy(),
        rmse=float(res.rmse[0].item()), fitness=inl / float(len(A)),
        indices=res.indices[0].cpu().numpy().astype(np.intp),
        src=res.src_final[0].cpu().numpy(),
    )


def _icp_full_large(A, B, max_iterations, tolerance, init_pose, max_corr_dist) -> IcpOutput:
    """Point sets beyond the fused per-pair kernel (> 1,024 source or > 4,096 target points, e.g. a
    scan against the 11 k-point local map of the reference's SLAM loop): the sharded-map path with
    a single shard -- same loop, same exactness, any size."""
    from .scan_to_map import MapShard, ScanToMap
    s2m = ScanToMap(MapShard(torch.from_numpy(B).cuda()), len(A), want_indices=True, local_only=True)
    ip = None
    if init_pose is not None:
        ip = _pose6(init_pose)[0].cpu().numpy()
    r = s2m.run(torch.from_numpy(A).cuda(), max_iterations=max_iterations, tolerance=tolerance,
                init_pose=ip, max_corr_dist=max_corr_dist)
    return IcpOutput(R=r.R, t=r.t, error=r.error, iterations=r.iterations, R_last=r.R_last, t_last=r.t_last,
                     rmse=r.rmse, fitness=r.inliers / float(len(A)),
                     indices=r.indices.cpu().numpy().astype(np.intp), src=r.src.cpu().numpy())


def icp(A, B, max_iterations: int = 20, tolerance: float = 1e-5, *, init_pose=None,
        max_corr_dist: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``icp(A, B, max_iterations=20, tolerance=1e-5) -> (src, R, t)``.

    Positional-compatible with labels_segmentation/icp.py:28.  Like the reference, the
    returned ``R, t`` are the LAST incremental update (icp.py:42,53), not the cumulative
    pose; use :func:`icp_full` for the cumulative pose, error and iteration count.
    """
    o = icp_full(A, B, max_iterations, tolerance, init_pose=init_pose, max_corr_dist=max_corr_dist)
    return o.src, o.R_last, o.t_last


def best_fit_transform(A, B) -> Tuple[np.ndarray, np.ndarray]:
    """``best_fit_transform(A, B) -> (R, t)`` for matched rows (labels_segmentation/icp.py:5-26)."""
    A = _as_points(A, "A")
    B = _as_points(B, "B")
    if len(A) != len(B):
        raise ValueError("A and B must have the same number of rows")     # NumPy would fail to broadcast
    pose = best_fit(ScanTable(torch.from_numpy(A[None]).cuda()), ScanTable(torch.from_numpy(B[None]).cuda()),
                    n_pairs=1)[0].cpu().numpy()
    return pose[:4].reshape(2, 2).copy(), pose[4:6].copy()


def nearest_neighbors(src, tgt) -> Tuple[np.ndarray, np.ndarray]:
    """``distances, indices = KDTree(tgt).query(src)`` (icp.py:37-38) on the device."""
    A = _as_points(src, "src")
    B = _as_points(tgt, "tgt")
    s = ScanTable(torch.from_numpy(A[None]).cuda(), None)
    t = ScanTable(torch.from_numpy(B[None]).cuda(), None)
    idx, d2 = nn_search(s, t, n_pairs=1)
    return np.sqrt(d2[0].cpu().numpy()), idx[0].cpu().numpy().astype(np.intp)


def registration_p2p(points1, points2, threshold: float = 200.0, voxel_size: Optional[float] = None,
                     trans_init=None, max_iteration: int = 50, tolerance: float = 1e-5):
    """Open3D-shaped point-to-point adapter: ``(rmse, T4x4)``.

    Positional-compatible with ``gicp(points1, points2, threshold, voxel_size, trans_init)``.
    Mirrors the call shape of ``gicp(points1, points2, threshold, voxel, trans_init)``
    (duc/ICP_LIDAR/gicp_lidar.py:12-36; caller duc/ICP_LIDAR/mainn.py:311) and of
    ``icp(...)`` in duc/code python/b.py:219-236, including the ``< 10`` points guard
    (gicp_lidar.py:13-15).  Gate / rmse semantics are parity-unpinned (Open3D is not
    vendored by the reference) and are defined by oracle.icp_oracle.icp_extended.
    ``points1`` / ``points2`` may be NumPy arrays or CUDA tensors ((N, 2) or (N, 3)); tensors stay
    on the device from the voxel grid to the pose (the SLAM loop passes its map that way).
    """
    if len(points1) < 10 or len(points2) < 10:
        return float("inf"), np.eye(4)

    def dev(p):
        if isinstance(p, torch.Tensor):
            t = p[:, :2].to(device="cuda", dtype=torch.float64)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(p, dtype=np.float64)[:, :2])).cuda()
        return t.contiguous()

    p1, p2 = dev(points1), dev(points2)
    if voxel_size:                       # gicp_lidar.py:20-21: both clouds are down-sampled first
        from .mapping import voxel_down_sample
        p1, p2 = voxel_down_sample(p1, voxel_size), voxel_down_sample(p2, voxel_size)
    lib = _cabi.lib()
    if p1.shape[0] > lib.b200icp_max_src_pitch() or p2.shape[0] > lib.b200icp_max_tgt_pitch():
        o = _icp_full_large(p1.cpu().numpy(), p2.cpu().numpy(), max_iteration, tolerance,
                            None if trans_init is None else np.asarray(trans_init), threshold)
        rmse, R, t = o.rmse, o.R, o.t
    else:
        res = align_pairs(ScanTable(p1[None].contiguous()), ScanTable(p2[None].contiguous()), n_pairs=1,
                          max_iterations=max_iteration, tolerance=tolerance,
                          init_pose=None if trans_init is None else _pose6(np.asarray(trans_init)),
                          max_corr_dist=threshold)
        host = torch.cat([res.pose_total[0], res.rmse]).cpu().numpy()       # one device-to-host read
        R, t, rmse = host[:4].reshape(2, 2), host[4:6], float(host[6])
    T = np.eye(4)
    T[:2, :2] = R
    T[:2, 3] = t
    return rmse, T
