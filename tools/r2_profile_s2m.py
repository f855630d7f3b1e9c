"""One scan-to-map alignment on the configs[4] instance (for ncu / timing)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m                       # noqa: E402
from oracle import icp_oracle as orc                 # noqa: E402  (synthetic inputs only)
M, N = int(os.environ.get("MAP", str(1 << 24))), int(os.environ.get("SCAN", "8192"))
shard = m.MapShard(torch.from_numpy(orc.synth_map(M)).cuda())
scan = torch.from_numpy(orc.synth_scan_for_map(N)).cuda()
s2m = m.ScanToMap(shard, N)
for _ in range(int(os.environ.get("REPS", "3"))):
    s2m.run(scan, max_iterations=30, tolerance=-1.0, sync=False)
torch.cuda.synchronize()
evs = []
s2m.run(scan, max_iterations=30, tolerance=-1.0, sync=False, events=evs)
torch.cuda.synchronize()
se = [evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(30)]
print("search ms per iteration:", " ".join(f"{x:.3f}" for x in se))
print(f"total {evs[0].elapsed_time(evs[-1]):.3f} ms; error {s2m.result().error:.9f}")
