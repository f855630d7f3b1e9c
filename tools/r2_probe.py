"""Round-2 probe (run under gpurun): numbers that decide the kernel work.
  * configs[2] with tolerance 1e-5: iteration histogram and launch time (current kernel);
  * scan-to-map: per-iteration error / increment (how many iterations really move the scan)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m                       # noqa: E402
from oracle import icp_oracle as orc                 # noqa: E402  (synthetic inputs only)


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def pairs():
    P = 65536
    src, tgt = orc.synth_room_batch(0, P)
    s, t = m.ScanTable(torch.from_numpy(src).cuda()), m.ScanTable(torch.from_numpy(tgt).cuda())
    out = m.alloc_outputs(P, 360, "cuda")
    for tol in (-1.0, 1e-5):
        ms = timed(lambda: m.align_pairs(s, t, max_iterations=30, tolerance=tol, out=out))
        it = out.iterations.cpu().numpy()
        print(f"pairs tol={tol}: {ms:.3f} ms; iterations mean {it.mean():.2f} min {it.min()} max {it.max()} "
              f"hist {np.bincount(it, minlength=31).tolist()}", flush=True)


def s2m(M=1 << 24, N=8192):
    full = orc.synth_map(M)
    shard = m.MapShard(torch.from_numpy(full).cuda())
    scan = torch.from_numpy(orc.synth_scan_for_map(N)).cuda()
    from icp_slam_yolo_b200.scan_to_map import ScanToMapLocalShards
    loc = ScanToMapLocalShards([shard], N)
    loc.init(scan)
    prev = None
    for it in range(30):
        loc.step(30, -1.0)
        torch.cuda.synchronize()
        st = loc.state.cpu().numpy()
        src = loc.src64.cpu().numpy()
        mv = 0.0 if prev is None else float(np.max(np.hypot(*(src - prev).T)))
        prev = src
        print(f"s2m it {it}: error {st[12]:.9f} inc theta {np.arctan2(st[8], st[6]):.3e} t ({st[10]:.3e},{st[11]:.3e}) max move {mv:.3e}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "pairs"):
        pairs()
    if which in ("all", "s2m"):
        s2m()
