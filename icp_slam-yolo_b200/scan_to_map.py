"""Scan-to-map ICP with the map sharded across GPUs (one process per GPU).

Call shape of the reference's SLAM step -- register the current scan against the (local) map
(duc/ICP_LIDAR/mainn.py:297-318, slam_offline.py:366-392) -- with the point-to-point loop of
labels_segmentation/icp.py:28-53.  The map is split into contiguous index ranges, one per
rank; per iteration every rank finds the exact nearest map point of ITS shard for every scan
point (CUDA), the 32-byte records are all-gathered (NCCL over NVLink; the only collective),
and every rank runs the same update kernel on the same gathered records, so poses, errors and
the stop decision are bit-identical everywhere without a second collective.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _cabi
from .registration import _DTYPES, _ptr, _require_cuda, _stream_ptr

RECORD_BYTES = 32
STATE_BYTES = 136


def _lib():
    return _cabi.lib()


class MapShard:
    """One rank's contiguous slice of the map, prepared for the sweep kernel."""

    def __init__(self, points: torch.Tensor, global_offset: int = 0, stream=None):
        if points.dim() != 2 or points.shape[1] != 2 or points.dtype not in _DTYPES:
            raise ValueError("map points must be [m, 2] float32/float64")
        _require_cuda(points, "map points")
        if points.shape[0] < 1:
            raise ValueError("empty map shard")
        self.points = points
        self.m = int(points.shape[0])
        self.global_offset = int(global_offset)
        chunk = _lib().b200icp_s2m_chunk()
        n_chunks = (self.m + chunk - 1) // chunk
        dev = points.device
        self.cx = torch.empty(n_chunks * chunk, dtype=torch.float32, device=dev)
        self.cy = torch.empty(n_chunks * chunk, dtype=torch.float32, device=dev)
        self.chunk_origin = torch.empty((n_chunks, 2), dtype=torch.float64, device=dev)
        self.chunk_radius = torch.empty(n_chunks, dtype=torch.float32, device=dev)
        self.desc = _cabi.S2MShard()
        self.desc.points = points.data_ptr()
        self.desc.m = self.m
        self.desc.global_offset = self.global_offset
        self.desc.dtype = _DTYPES[points.dtype]
        self.desc.cx, self.desc.cy = self.cx.data_ptr(), self.cy.data_ptr()
        self.desc.chunk_origin = self.chunk_origin.data_ptr()
        self.desc.chunk_radius = self.chunk_radius.data_ptr()
        with torch.cuda.device(dev):
            rc = _lib().b200icp_s2m_prepare_map(C.byref(self.desc), _stream_ptr(stream))
        _cabi.check(rc, "b200icp_s2m_prepare_map")


@dataclass
class ScanToMapResult:
    R: np.ndarray             # cumulative rotation (2,2)
    t: np.ndarray             # cumulative translation (2,)
    R_last: np.ndarray
    t_last: np.ndarray
    error: float              # lagged mean NN distance (icp.py:48)
    rmse: float
    inliers: int
    iterations: int
    indices: Optional[torch.Tensor]      # [n] int32 global map index of the last search
    src: torch.Tensor                    # [n,2] float64 transformed scan (device)


class PeerExchange:
    """Peer-visible record buffers of all ranks (CUDA IPC), for the store-to-every-peer all-gather
    (b200icp_s2m_publish / b200icp_s2m_wait).  Layout per rank: [2 slots][world][n] records, then
    [2][world] int64 flags.  NCCL is used once, to exchange the 64-byte IPC handles."""

    def __init__(self, n: int, world: int, rank: int, group, device):
        import torch.distributed as dist
        self.n, self.world, self.rank, self.device = n, world, rank, device
        self.slot_bytes = world * n * RECORD_BYTES
        nbytes = 2 * self.slot_bytes + 2 * world * 8
        lib = _lib()
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(device):
            _cabi.check(lib.b200icp_peer_alloc(nbytes, C.byref(ptr), handle), "b200icp_peer_alloc")
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(device)
            everyone = torch.empty(world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(everyone, mine, group=group)
            handles = everyone.cpu().numpy().reshape(world, 64)
            self.base = int(ptr.value)
            self.opened = []
            addrs = []
            for r in range(world):
                if r == rank:
                    addrs.append(self.base)
                    continue
                p = C.c_void_p()
                _cabi.check(lib.b200icp_peer_open(handles[r].tobytes(), C.byref(p)), "b200icp_peer_open")
                self.opened.append(int(p.value))
                addrs.append(int(p.value))
            self.peers = torch.tensor(addrs, dtype=torch.int64, device=device)
            self.counter = torch.zeros(4, dtype=torch.int32, device=device)
            self.seq = 0
            torch.cuda.synchronize()
            dist.barrier(group=group)

    def slot_ptr(self, slot: int) -> C.c_void_p:
        return C.c_void_p(self.base + slot * self.slot_bytes)

    def close(self):
        import torch.distributed as dist
        lib = _lib()
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        for p in self.opened:
            lib.b200icp_peer_close(C.c_void_p(p))
        self.opened = []
        if self.base:
            lib.b200icp_peer_free(C.c_void_p(self.base))
            self.base = 0


class ScanToMap:
    """Reusable buffers for registering n-point scans against one shard (per rank).

    exchange = "nccl": records all-gathered with torch.distributed (default);
    exchange = "peer": every rank stores its records straight into every peer's buffer over
    NVLink and raises a flag there (PeerExchange) -- no library collective in the loop."""

    def __init__(self, shard: MapShard, n_scan: int, group=None, want_indices: bool = False,
                 exchange: str = "nccl", local_only: bool = False):
        import torch.distributed as dist
        self.shard, self.n = shard, int(n_scan)
        self.group = group
        self.world = (dist.get_world_size(group)
                      if (not local_only and dist.is_available() and dist.is_initialized()) else 1)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.exchange = exchange if self.world > 1 else "nccl"
        self.peer = None
        dev = shard.points.device
        self.dev = dev
        self.src64 = torch.empty((self.n, 2), dtype=torch.float64, device=dev)
        self.state = torch.zeros(STATE_BYTES // 8, dtype=torch.float64, device=dev)
        self.records = torch.empty((self.n, 4), dtype=torch.float64, device=dev)          # 32 B each
        self.ub = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.records_all = (torch.empty((self.world, self.n, 4), dtype=torch.float64, device=dev)
                            if self.world > 1 else None)
        wb = _lib().b200icp_s2m_workspace_bytes(self.n, shard.m)
        if wb < 0:
            raise _cabi.B200IcpError("b200icp_s2m_workspace_bytes failed")
        self.workspace = torch.empty(wb, dtype=torch.uint8, device=dev)
        self.indices = torch.empty(self.n, dtype=torch.int32, device=dev) if want_indices else None
        self.launches = 0
        if self.exchange == "peer":
            self.peer = PeerExchange(self.n, self.world, self.rank, group, dev)
        elif self.exchange != "nccl":
            raise ValueError("exchange must be 'nccl' or 'peer'")

    def search(self, stream=None):
        """records <- exact nearest point of this shard for the current scan state (or "none" for
        points whose nearest neighbour is provably in another rank's shard)."""
        import torch.distributed as dist
        rc = _lib().b200icp_s2m_bound(C.byref(self.shard.desc), _ptr(self.src64), self.n, _ptr(self.ub),
                                      _ptr(self.state), _stream_ptr(stream))
        _cabi.check(rc, "b200icp_s2m_bound")
        if self.world > 1:             # global bound: 4 bytes per scan point
            dist.all_reduce(self.ub, op=dist.ReduceOp.MIN, group=self.group)
        rc = _lib().b200icp_s2m_search(C.byref(self.shard.desc), _ptr(self.src64), self.n, _ptr(self.ub),
                                       _ptr(self.records), _ptr(self.workspace),
                                       self.workspace.numel(), _ptr(self.state), _stream_ptr(stream))
        _cabi.check(rc, "b200icp_s2m_search")
        self.launches += 6

    def run(self, scan: torch.Tensor, *, max_iterations: int = 20, tolerance: float = 1e-5,
            init_pose=None, max_corr_dist: Optional[float] = None, sync: bool = True):
        """Full loop.  No host synchronisation inside; kernels no-op once converged."""
        import torch.distributed as dist
        if scan.dim() != 2 or scan.shape != (self.n, 2) or scan.dtype not in _DTYPES:
            raise ValueError(f"scan must be [{self.n}, 2] float32/float64")
        _require_cuda(scan, "scan")
        ip = None
        if init_pose is not None:
            ip = torch.as_tensor(np.asarray(init_pose, dtype=np.float64).reshape(6)).to(self.dev)
        self.launches = 0
        with torch.cuda.device(self.dev):
            rc = _lib().b200icp_s2m_init(_ptr(scan), _DTYPES[scan.dtype], self.n, _ptr(ip),
                                         _ptr(self.src64), _ptr(self.state), _stream_ptr(None))
            _cabi.check(rc, "b200icp_s2m_init")
            self.launches += 1
            for _ in range(int(max_iterations)):
                self.search()
                if self.world > 1 and self.peer is not None:
                    pe = self.peer
                    pe.seq += 1
                    slot = pe.seq & 1
                    rc = _lib().b200icp_s2m_publish(_ptr(self.records), self.n, _ptr(pe.peers), self.world,
                                                    self.rank, slot, pe.seq, _ptr(pe.counter),
                                                    _ptr(self.state), _stream_ptr(None))
                    _cabi.check(rc, "b200icp_s2m_publish")
                    rc = _lib().b200icp_s2m_wait(C.c_void_p(pe.base), self.n, self.world, slot, pe.seq,
                                                 _ptr(self.state), _stream_ptr(None))
                    _cabi.check(rc, "b200icp_s2m_wait")
                    self.launches += 2
                    rec_all, ranks = pe.slot_ptr(slot), self.world
                elif self.world > 1:
                    dist.all_gather_into_tensor(self.records_all.view(self.world * self.n, 4),
                                                self.records, group=self.group)
                    rec_all, ranks = _ptr(self.records_all), self.world
                else:
                    rec_all, ranks = _ptr(self.records), 1
                rc = _lib().b200icp_s2m_update(rec_all, ranks, _ptr(self.src64), self.n,
                                               int(max_iterations), float(tolerance),
                                               0.0 if max_corr_dist is None else float(max_corr_dist),
                                               _ptr(self.indices), _ptr(self.state), _stream_ptr(None))
                _cabi.check(rc, "b200icp_s2m_update")
                self.launches += 1
        return self.result() if sync else None

    def result(self) -> ScanToMapResult:
        st = self.state.cpu().numpy()
        ints = st[15:17].view(np.int32)
        mean_d2 = st[13]
        return ScanToMapResult(
            R=st[0:4].reshape(2, 2).copy(), t=st[4:6].copy(),
            R_last=st[6:10].reshape(2, 2).copy(), t_last=st[10:12].copy(),
            error=float(st[12]), rmse=float(np.sqrt(mean_d2)) if np.isfinite(mean_d2) else float("inf"),
            inliers=int(ints[1]), iterations=int(ints[0]), indices=self.indices, src=self.src64)


def scan_to_map_icp(scan: torch.Tensor, shard: MapShard, max_iterations: int = 20,
                    tolerance: float = 1e-5, *, init_pose=None, max_corr_dist=None, group=None,
                    want_indices: bool = False) -> ScanToMapResult:
    """One-shot convenience wrapper around :class:`ScanToMap`."""
    s2m = ScanToMap(shard, int(scan.shape[0]), group=group, want_indices=want_indices)
    return s2m.run(scan, max_iterations=max_iterations, tolerance=tolerance, init_pose=init_pose,
                   max_corr_dist=max_corr_dist)


class ScanToMapLocalShards:
    """The multi-rank protocol with every shard resident on ONE GPU: each shard is searched in
    turn, the records are stacked where the all-gather would put them, and the same update
    kernel runs.  Used to split a map that is built incrementally into pieces, and by the tests
    to check the sharded path (shard offsets, record merge, lowest-global-index ties) without
    several GPUs.  ``step()`` runs one iteration so callers can inspect per-iteration state."""

    def __init__(self, shards, n_scan: int, want_indices: bool = True):
        self.workers = [ScanToMap(s, n_scan, local_only=True) for s in shards]
        self.n, self.dev = int(n_scan), shards[0].points.device
        w0 = self.workers[0]
        self.src64, self.state = w0.src64, w0.state
        self.records_all = torch.empty((len(shards), self.n, 4), dtype=torch.float64, device=self.dev)
        self.indices = torch.empty(self.n, dtype=torch.int32, device=self.dev) if want_indices else None

    def init(self, scan: torch.Tensor, init_pose=None):
        ip = None
        if init_pose is not None:
            ip = torch.as_tensor(np.asarray(init_pose, dtype=np.float64).reshape(6)).to(self.dev)
        rc = _lib().b200icp_s2m_init(_ptr(scan), _DTYPES[scan.dtype], self.n, _ptr(ip),
                                     _ptr(self.src64), _ptr(self.state), _stream_ptr(None))
        _cabi.check(rc, "b200icp_s2m_init")

    def step(self, max_iterations: int, tolerance: float, max_corr_dist=None):
        for w in self.workers:         # per-shard bounds, then the "all-reduce": elementwise min
            rc = _lib().b200icp_s2m_bound(C.byref(w.shard.desc), _ptr(self.src64), self.n, _ptr(w.ub),
                                          _ptr(self.state), _stream_ptr(None))
            _cabi.check(rc, "b200icp_s2m_bound")
        ub = self.workers[0].ub
        for w in self.workers[1:]:
            torch.minimum(ub, w.ub, out=ub)
        for g, w in enumerate(self.workers):
            rec = self.records_all[g]
            rc = _lib().b200icp_s2m_search(C.byref(w.shard.desc), _ptr(self.src64), self.n, _ptr(ub),
                                           _ptr(rec), _ptr(w.workspace), w.workspace.numel(),
                                           _ptr(self.state), _stream_ptr(None))
            _cabi.check(rc, "b200icp_s2m_search")
        rc = _lib().b200icp_s2m_update(_ptr(self.records_all), len(self.workers), _ptr(self.src64),
                                       self.n, int(max_iterations), float(tolerance),
                                       0.0 if max_corr_dist is None else float(max_corr_dist),
                                       _ptr(self.indices), _ptr(self.state), _stream_ptr(None))
        _cabi.check(rc, "b200icp_s2m_update")

    def run(self, scan, *, max_iterations=20, tolerance=1e-5, init_pose=None, max_corr_dist=None):
        self.init(scan, init_pose)
        for _ in range(int(max_iterations)):
            self.step(max_iterations, tolerance, max_corr_dist)
        return self.result()

    def result(self) -> ScanToMapResult:
        w0 = self.workers[0]
        w0.indices = self.indices
        return w0.result()
