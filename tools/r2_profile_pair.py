"""One launch of the throughput kernel on the configs[2] batch (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m                       # noqa: E402
from oracle import icp_oracle as orc                 # noqa: E402  (synthetic inputs only)
P = int(os.environ.get("PAIRS", "65536"))
W = int(os.environ.get("PAIR_WARPS", "0"))
TOL = float(os.environ.get("TOL", "-1"))
KERNEL = os.environ.get("KERNEL", "warp")
src, tgt = orc.synth_room_batch(0, P)
s, t = m.ScanTable(torch.from_numpy(src).cuda()), m.ScanTable(torch.from_numpy(tgt).cuda())
out = m.alloc_outputs(P, 360, "cuda")
for _ in range(int(os.environ.get("REPS", "2"))):
    m.align_pairs(s, t, max_iterations=30, tolerance=TOL, kernel=KERNEL, pair_warps=W, out=out,
                  dense_sweep=os.environ.get("DENSE", "0") == "1", sweep_reuse=os.environ.get("REUSE", "1") == "1")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.align_pairs(s, t, max_iterations=30, tolerance=TOL, kernel=KERNEL, pair_warps=W, out=out,
              dense_sweep=os.environ.get("DENSE", "0") == "1", sweep_reuse=os.environ.get("REUSE", "1") == "1")
e1.record()
torch.cuda.synchronize()
print(f"P={P} W={W} tol={TOL} kernel={KERNEL}: {e0.elapsed_time(e1):.3f} ms")
