import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m
from icp_slam_yolo_b200 import scan_io
raw = scan_io.unpack_fixture(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scan_data_1_packed.npz"))
n = int(os.environ.get("NSCANS", "64"))
table = scan_io.prepare_scans(raw[900:900 + n], device="cuda")
kw = dict(max_iterations=30, tolerance=1e-5, want_indices=True, want_stats=True, want_history=True)
a = m.align_consecutive(table, **kw)
a2 = m.align_consecutive(table, **kw)
d = m.align_consecutive(table, dense_sweep=True, **kw)
d2 = m.align_consecutive(table, dense_sweep=True, **kw)
print("pruned deterministic", torch.equal(a.pose_total, a2.pose_total), "dense deterministic", torch.equal(d.pose_total, d2.pose_total))
print("iterations equal", torch.equal(a.iterations, d.iterations), "indices equal", torch.equal(a.indices, d.indices),
      "history equal", torch.equal(a.index_history, d.index_history))
print("max pose diff", float((a.pose_total - d.pose_total).abs().max()), "error diff", float((a.error - d.error).abs().max()))
bad = (a.pose_total != d.pose_total).any(dim=1).nonzero().flatten().tolist()
print("pairs differing:", bad[:20], "lengths", table.lengths[[b + 1 for b in bad[:10]]].tolist() if bad else None)
