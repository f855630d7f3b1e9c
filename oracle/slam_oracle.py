"""CPU oracle of the offline SLAM loop: the composition of the oracle restatements in the order
of duc/ICP_LIDAR/slam_offline.py:318-455 (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Every step is one of the pinned oracles -- scan-to-map registration with gate and initial pose
(oracle.icp_oracle.icp_extended behind the gicp() call shape, slam_offline.py:382), 2D voxel grid,
dynamic-point removal by NN distance (process.py:75-84), occupancy filter / ray casting / pruning
(oracle.occupancy_oracle).  The loop as a whole is parity-UNPINNED against the reference, whose
registration step is Open3D's Generalized ICP (not vendored); it defines what
icp_slam-yolo_b200/slam.py must reproduce."""
import numpy as np

from . import icp_oracle as orc
from . import occupancy_oracle as occ


class OracleSlam:
    """slam_offline.py:318-455 with the oracle restatements (NumPy / SciPy / plain C)."""

    def __init__(self, cfg):
        self.c = cfg
        self.center = (cfg.map_width_pixels // 2, cfg.map_height_pixels // 2)
        self.occ = np.full((cfg.map_height_pixels, cfg.map_width_pixels), 0.5, dtype=np.float32)
        self.img = np.full((cfg.map_height_pixels, cfg.map_width_pixels, 3), 128, dtype=np.uint8)
        self.map = np.zeros((0, 2))
        self.pose = np.eye(4)
        self.prev = None
        self.cur = np.zeros((0, 3))
        self.mapped = False

    def _occupancy(self):
        if len(self.cur):
            occ.update_occupancy_map_c(self.occ, self.img, self.cur, self.pose[:3, 3], self.center,
                                       self.c.resolution_mm_per_pixel)
            self.mapped = True

    def first(self, pts):
        self.map = np.ascontiguousarray(pts[:, :2])
        self.cur = pts
        self._occupancy()

    def _gicp(self, p1, p2):
        c = self.c
        if len(p1) < 10 or len(p2) < 10:
            return float("inf"), np.eye(4)
        a, b = orc.voxel_down_sample_2d(p1[:, :2], c.icp_voxel_size), orc.voxel_down_sample_2d(p2[:, :2], c.icp_voxel_size)
        o = orc.icp_extended(a, b, c.max_iteration, c.tolerance, init_pose=(self.pose[:2, :2], self.pose[:2, 3]),
                             max_corr_dist=c.icp_threshold)
        T = np.eye(4)
        T[:2, :2], T[:2, 3] = o.R_tot, o.t_tot
        return o.rmse, T

    def step(self, pts):
        c = self.c
        if pts is None or len(pts) < c.min_scan_points:
            return None
        robot = self.pose[:3, 3]
        local = self.map
        if len(self.map):
            keep = np.sum((self.map - robot[:2]) ** 2, axis=1) < c.local_map_radius_mm ** 2
            local = self.map[keep] if keep.sum() >= c.min_icp_map_points else self.map
        rmse, T = self._gicp(pts, local)
        if rmse > c.max_rmse_threshold:
            return False, rmse
        self.pose = T
        self.cur = np.dot(pts, T[:3, :3].T) + T[:3, 3]
        add = self.cur[:, :2]
        if self.prev is not None and len(self.prev):
            d, _ = orc.nn_bruteforce(add, self.prev)                   # process.py:80-82: distances < threshold
            add = add[d < c.dynamic_distance_threshold]
        if self.mapped and len(add):
            add = add[occ.filter_points_by_occupancy(add, self.occ, self.center, c.resolution_mm_per_pixel)]
        if len(add):
            self.map = np.concatenate([self.map, add])
        if len(self.map) > c.max_map_points_before_downsample:
            self.map = orc.voxel_down_sample_2d(self.map, c.icp_voxel_size)
        self.prev = self.cur[:, :2].copy()
        self._occupancy()
        if self.mapped and len(self.map):
            self.map = self.map[occ.filter_points_by_occupancy(self.map, self.occ, self.center, c.resolution_mm_per_pixel)]
        return True, rmse
