"""Sequence odometry over a scan table: every consecutive pair in ONE launch.

The reference's SLAM loops are sequential because each frame's pose seeds the next
registration (duc/ICP_LIDAR/slam_offline.py:344-392).  ``icp()`` itself has no warm start
(labels_segmentation/icp.py:28), so pairwise alignments (scan k+1 -> scan k) are
independent; the global pose is the prefix composition T_{0,k+1} = T_{0,k} o T_k, an O(B)
scan over 6 numbers per pair done after the batched launch.
"""
from __future__ import annotations

import numpy as np
import torch

from .registration import AlignResult, ScanTable, align_pairs


def align_consecutive(table: ScanTable, *, max_iterations: int = 30, tolerance: float = 1e-5,
                      max_corr_dist=None, out=None, **kw) -> AlignResult:
    """Pair p aligns row p+1 (source) onto row p (target)."""
    return align_pairs(table.slice_rows(1), table.slice_rows(0, table.rows - 1),
                       n_pairs=table.rows - 1, max_iterations=max_iterations,
                       tolerance=tolerance, max_corr_dist=max_corr_dist, out=out, **kw)


def chain_poses_device(pose_total: torch.Tensor, stream=None) -> torch.Tensor:
    """Prefix-compose pairwise poses [B,6] (CUDA, float64) into global poses [B+1,6] on the device."""
    import ctypes as C
    from . import _cabi
    from .registration import _ptr, _require_cuda, _stream_ptr
    if pose_total.dtype != torch.float64 or pose_total.dim() != 2 or pose_total.shape[1] != 6:
        raise ValueError("pose_total must be float64 [B, 6]")
    _require_cuda(pose_total, "pose_total")
    out = torch.empty((pose_total.shape[0] + 1, 6), dtype=torch.float64, device=pose_total.device)
    with torch.cuda.device(pose_total.device):
        rc = _cabi.lib().b200icp_chain_poses(_ptr(pose_total), int(pose_total.shape[0]), _ptr(out), _stream_ptr(stream))
    _cabi.check(rc, "b200icp_chain_poses")
    return out


def chain_poses(pose_total) -> np.ndarray:
    """Prefix-compose pairwise poses [B,6] into global poses [B+1,6] (row 0 = identity).
    CUDA tensors are composed on the device (one kernel); arrays on the host."""
    if isinstance(pose_total, torch.Tensor) and pose_total.is_cuda:
        return chain_poses_device(pose_total).cpu().numpy()
    p = pose_total.detach().cpu().numpy() if isinstance(pose_total, torch.Tensor) else np.asarray(pose_total)
    out = np.zeros((len(p) + 1, 6))
    R, t = np.eye(2), np.zeros(2)
    out[0] = [1, 0, 0, 1, 0, 0]
    for k in range(len(p)):
        Rk, tk = p[k, :4].reshape(2, 2), p[k, 4:6]
        t = R @ tk + t
        R = R @ Rk
        out[k + 1, :4] = R.reshape(4)
        out[k + 1, 4:6] = t
    return out
