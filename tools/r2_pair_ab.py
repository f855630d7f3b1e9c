"""A/B of the W-warps-per-pair fused kernel against the round-1 warp kernel (run under gpurun):
bit-identical index histories on synthetic rooms and real scans, then launch times."""
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


def check(s, t, n_pairs, tol, label, **kw):
    base = dict(n_pairs=n_pairs, max_iterations=30, tolerance=tol, want_history=True, want_stats=True)
    ref = m.align_pairs(s, t, kernel="legacy-warp", **base)
    ok = True
    for w in (0, 1, 2, 3, 4):
        for dense in (False, True):
            for reuse in (True, False):
                new = m.align_pairs(s, t, kernel="warp", pair_warps=w, dense_sweep=dense, sweep_reuse=reuse, **base)
                same_h = torch.equal(new.index_history, ref.index_history)
                same_i = torch.equal(new.iterations, ref.iterations)
                dp = float((new.pose_total - ref.pose_total).abs().max())
                de = float((new.error - ref.error).abs().max())
                if not (same_h and same_i and dp < 1e-9 and de < 1e-9):
                    ok = False
                    bad = int((new.index_history != ref.index_history).any(dim=2).any(dim=1).sum())
                    print(f"  MISMATCH {label} W={w} dense={dense} reuse={reuse}: hist {same_h} ({bad} pairs) iters {same_i} dpose {dp:.2e} derr {de:.2e}")
                elif w == 0:
                    print(f"  ok {label} W=auto dense={dense} reuse={reuse}: dpose {dp:.2e}; evals new {int(new.evaluated_pairs.sum())} ref {int(ref.evaluated_pairs.sum())}")
    return ok


def main():
    P = 2048
    src, tgt = orc.synth_room_batch(0, P)
    s, t = m.ScanTable(torch.from_numpy(src).cuda()), m.ScanTable(torch.from_numpy(tgt).cuda())
    ok = check(s, t, P, -1.0, "rooms tol=-1")
    ok &= check(s, t, P, 1e-5, "rooms tol=1e-5")
    from icp_slam_yolo_b200 import scan_io
    raw = scan_io.unpack_fixture(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scan_data_1_packed.npz"))
    table = scan_io.prepare_scans(raw, device="cuda")
    ok &= check(table.slice_rows(1), table.slice_rows(0, table.rows - 1), table.rows - 1, 1e-5, "scan_data_1")
    print("ALL OK" if ok else "FAILURES", flush=True)
    P = 65536
    src, tgt = orc.synth_room_batch(0, P)
    s, t = m.ScanTable(torch.from_numpy(src).cuda()), m.ScanTable(torch.from_numpy(tgt).cuda())
    out = m.alloc_outputs(P, 360, "cuda")
    for tol in (-1.0, 1e-5):
        ms = timed(lambda: m.align_pairs(s, t, max_iterations=30, tolerance=tol, kernel="legacy-warp", out=out))
        print(f"legacy tol={tol}: {ms:.3f} ms", flush=True)
        for w in (1, 2, 3, 4):
            ms = timed(lambda: m.align_pairs(s, t, max_iterations=30, tolerance=tol, kernel="warp", pair_warps=w, out=out))
            print(f"pair W={w} tol={tol}: {ms:.3f} ms", flush=True)
    for w in (2, 3, 4):
        ms = timed(lambda: m.align_pairs(s, t, max_iterations=30, tolerance=-1.0, kernel="warp", pair_warps=w, dense_sweep=True, sweep_reuse=False, out=out), reps=3)
        print(f"pair dense W={w}: {ms:.3f} ms", flush=True)
    ms = timed(lambda: m.align_pairs(s, t, max_iterations=30, tolerance=-1.0, kernel="legacy-warp", dense_sweep=True, sweep_reuse=False, out=out), reps=3)
    print(f"legacy dense: {ms:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
