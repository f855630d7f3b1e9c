"""Run under torchrun on N GPUs: scan-to-map ICP with the map sharded over the ranks must give
bit-identical state on every rank, and the same indices/pose as the oracle at reduced size.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/check_s2m_multigpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m          # noqa: E402
from oracle import icp_oracle as orc    # noqa: E402  (checker only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    M, N, iters = 1 << 18, 8192, 6
    full = orc.synth_map(M)
    scan_np = orc.synth_scan_for_map(N)
    b, e = m.shard_range(M, rank, world)
    shard = m.MapShard(torch.from_numpy(full[b:e]).to(dev), global_offset=b)
    exchange = os.environ.get("B200ICP_S2M_EXCHANGE", "nccl")
    s2m = m.ScanToMap(shard, N, want_indices=True, exchange=exchange)
    res = s2m.run(torch.from_numpy(scan_np).to(dev), max_iterations=iters, tolerance=-1.0)
    res = s2m.run(torch.from_numpy(scan_np).to(dev), max_iterations=iters, tolerance=-1.0)   # buffers reusable
    state = s2m.state.clone()
    gathered = [torch.empty_like(state) for _ in range(world)]
    dist.all_gather(gathered, state)
    same = all(torch.equal(g.view(torch.int64), gathered[0].view(torch.int64)) for g in gathered)
    if rank == 0:
        o = orc.icp_extended(scan_np, full, iters, -1.0)
        ok_idx = np.array_equal(res.indices.cpu().numpy(), o.indices[-1])
        dR = float(np.max(np.abs(res.R - o.R_tot)))
        dt = float(np.max(np.abs(res.t - o.t_tot)))
        print(f"exchange={exchange} world={world} state bit-identical across ranks: {same}; last-iteration indices == oracle: {ok_idx}; "
              f"|dR|={dR:.2e} |dt|={dt:.2e} mm; error={res.error:.9f} (oracle {o.error:.9f})", flush=True)
        assert same and ok_idx and dR < 1e-9 and dt < 1e-6
    if s2m.peer is not None:
        s2m.peer.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
