"""Latency of ONE gated scan-to-local-map registration (the per-frame call of the reference's SLAM
loop, mainn.py:311) through the three device paths: warp-per-pair kernel, CTA-per-pair kernel,
sharded-map path with one shard.  Run under gpurun."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m                       # noqa: E402
from oracle import icp_oracle as orc                 # noqa: E402  (synthetic inputs only)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    full = orc.synth_map(1 << 16, dtype=np.float64)
    scan = orc.synth_scan_for_map(160, dtype=np.float64)
    for mtgt in (1000, 3000):
        idx = np.sort(np.argsort(np.sum((full - scan.mean(0)) ** 2, axis=1))[:mtgt])     # the nearest map points, map order
        tgt = np.ascontiguousarray(full[idx])
        s = m.ScanTable(torch.from_numpy(scan[None]).cuda())
        t = m.ScanTable(torch.from_numpy(tgt[None]).cuda())
        out = m.alloc_outputs(1, s.pitch, "cuda")
        kw = dict(n_pairs=1, max_iterations=50, tolerance=1e-5, max_corr_dist=180.0, out=out)
        w = timed(lambda: m.align_pairs(s, t, kernel="warp", **kw))
        its = int(out.iterations[0].item())
        pw = out.pose_total.clone()
        b = timed(lambda: m.align_pairs(s, t, kernel="cta", **kw))
        same = bool(torch.allclose(pw, out.pose_total, atol=1e-9))
        shard = m.MapShard(torch.from_numpy(tgt).cuda())
        s2m = m.ScanToMap(shard, len(scan), local_only=True)
        sc = torch.from_numpy(scan).cuda()
        g = timed(lambda: s2m.run(sc, max_iterations=50, tolerance=1e-5, max_corr_dist=180.0), reps=10)
        print(f"160 x {mtgt}: iterations {its}; warp kernel {w:.3f} ms, CTA kernel {b:.3f} ms (same pose: {same}), "
              f"sharded-map path {g:.3f} ms")


def batches():
    """Where does the warp-per-pair kernel overtake the CTA-per-pair kernel?  360 x 360 rooms, tol 1e-5."""
    src, tgt = orc.synth_room_batch(9000, 4736)
    S, T = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
    for P in (16, 148, 592, 1184, 2368, 4736):
        s, t = m.ScanTable(S[:P].contiguous()), m.ScanTable(T[:P].contiguous())
        out = m.alloc_outputs(P, 360, "cuda")
        kw = dict(n_pairs=P, max_iterations=30, tolerance=1e-5, out=out)
        w = timed(lambda: m.align_pairs(s, t, kernel="warp", **kw), reps=10)
        b = timed(lambda: m.align_pairs(s, t, kernel="cta", **kw), reps=10)
        print(f"{P} pairs 360 x 360: warp kernel {w:.3f} ms, CTA kernel {b:.3f} ms")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "batches":
        batches()
        sys.exit(0)
    main()
