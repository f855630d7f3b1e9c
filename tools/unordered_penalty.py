"""How much slower is the pruned sweep than the dense one when the scans carry no spatial order
(points of every scan shuffled)?  Run under gpurun."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m
from oracle import icp_oracle as orc

P = 8192
src, tgt = orc.synth_room_batch(0, 256)
rng = np.random.default_rng(0)
for name, shuffle in (("ordered", False), ("shuffled", True)):
    s, t = src.copy(), tgt.copy()
    if shuffle:
        for b in range(len(s)):
            s[b] = s[b][rng.permutation(360)]
            t[b] = t[b][rng.permutation(360)]
    S = m.ScanTable(torch.from_numpy(s).cuda().repeat(P // 256, 1, 1))
    T = m.ScanTable(torch.from_numpy(t).cuda().repeat(P // 256, 1, 1))
    for mode in ("1", "0"):
        out = m.alloc_outputs(P, 360, "cuda", want_stats=True)
        for _ in range(2):
            m.align_pairs(S, T, max_iterations=30, tolerance=-1.0, dense_sweep=mode == "0", out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m.align_pairs(S, T, max_iterations=30, tolerance=-1.0, dense_sweep=mode == "0", out=out)
        e1.record(); torch.cuda.synchronize()
        frac = float(out.evaluated_pairs.sum().item()) / (P * 360 * 360 * 30)
        print(f"{name:9s} prune={mode}: {e0.elapsed_time(e1) / 3:7.3f} ms per {P} pairs, executed fraction {frac:.3f}")
