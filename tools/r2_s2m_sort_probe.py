"""Scan-to-map on the configs[4] instance with and without the Morton-sorted copy, and on a
shuffled copy of the same map (run under gpurun)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m
from oracle import icp_oracle as orc
M, N = 1 << 24, 8192
full = orc.synth_map(M)
scan = torch.from_numpy(orc.synth_scan_for_map(N)).cuda()
rng = np.random.default_rng(0)
for name, pts, sort in (("ordered, as given", full, False), ("ordered, Morton copy", full, True),
                        ("shuffled, Morton copy (auto)", full[rng.permutation(M)], "auto")):
    d = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    shard = m.MapShard(d, spatial_sort=sort)
    torch.cuda.synchronize(); prep = (time.perf_counter() - t0) * 1e3
    s2m = m.ScanToMap(shard, N)
    for _ in range(3):
        s2m.run(scan, max_iterations=30, tolerance=-1.0, sync=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        s2m.run(scan, max_iterations=30, tolerance=-1.0, sync=False)
    e1.record(); torch.cuda.synchronize()
    r = s2m.result()
    print(f"{name}: prepare {prep:.1f} ms, sorted={shard.order is not None}, {e0.elapsed_time(e1) / 5:.3f} ms per alignment, error {r.error:.9f}", flush=True)
    del s2m, shard, d
