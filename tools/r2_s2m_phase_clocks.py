"""Per-phase cycle split of s2m_search_kernel (needs a -DS2M_TIMING=1 variant: B200ICP_LIB=...)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m                       # noqa: E402
from icp_slam_yolo_b200 import _cabi                 # noqa: E402
from oracle import icp_oracle as orc                 # noqa: E402  (synthetic inputs only)
M, N = 1 << 24, 8192
shard = m.MapShard(torch.from_numpy(orc.synth_map(M)).cuda())
scan = torch.from_numpy(orc.synth_scan_for_map(N)).cuda()
s2m = m.ScanToMap(shard, N)
s2m.run(scan, max_iterations=30, tolerance=-1.0, sync=False)
torch.cuda.synchronize()
fn = _cabi.lib().b200icp_s2m_debug_clocks
fn.restype = C.c_int; fn.argtypes = [C.POINTER(C.c_ulonglong)]
buf = (C.c_ulonglong * 8)()
fn(buf)
for its in (8, 30):
    s2m.run(scan, max_iterations=its, tolerance=-1.0, sync=False)
    torch.cuda.synchronize()
    fn(buf)
    names = ["point+bound", "traversal", "scan", "record+store", "cta barrier"]
    tot = sum(buf[:5])
    print(f"{its} iterations: cycles per warp and iteration: " +
          ", ".join(f"{nm} {buf[k] / (N * its):.0f} ({100 * buf[k] / tot:.0f}%)" for k, nm in enumerate(names)))
