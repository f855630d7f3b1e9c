"""Run under torchrun on N GPUs: scan-to-map ICP with the map sharded over the ranks must give
bit-identical state on every rank, and the same indices / pose / error as the CPU oracle at
reduced size (M = 2^18), for the peer-store exchange, its CUDA-graph replay and the
NCCL exchange.  tests/test_multigpu.py runs this when at least two GPUs are visible.

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
    scan = torch.from_numpy(scan_np).to(dev)
    o = orc.icp_extended(scan_np, full, iters, -1.0) if rank == 0 else None
    o_tol = orc.icp_extended(scan_np, full, 30, 0.5) if rank == 0 else None
    ok_all = True
    for exchange, graph in (("peer", False), ("peer", True), ("nccl", False)):
        s2m = m.ScanToMap(shard, N, want_indices=True, exchange=exchange, graph=graph)
        for rep in range(3):                                                   # buffers / graph are reusable
            res = s2m.run(scan, max_iterations=iters, tolerance=-1.0)
        last_idx = res.indices.clone()
        state = s2m.state.clone()
        gathered = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(gathered, state)
        same = all(torch.equal(g.view(torch.int64), gathered[0].view(torch.int64)) for g in gathered)
        src_all = [torch.empty_like(s2m.src64) for _ in range(world)]
        dist.all_gather(src_all, s2m.src64)
        same_src = all(torch.equal(g.view(torch.int64), src_all[0].view(torch.int64)) for g in src_all)
        res_tol = s2m.run(scan, max_iterations=30, tolerance=0.5)              # early stop: same decision everywhere
        its = torch.tensor([res_tol.iterations], device=dev)
        its_all = [torch.empty_like(its) for _ in range(world)]
        dist.all_gather(its_all, its)
        same_its = len({int(t) for t in its_all}) == 1
        if rank == 0:
            ok_idx = np.array_equal(last_idx.cpu().numpy(), o.indices[-1])
            dR = float(np.max(np.abs(res.R - o.R_tot)))
            dt = float(np.max(np.abs(res.t - o.t_tot)))
            de = abs(res.error - o.error)
            ok = (same and same_src and same_its and ok_idx and dR < 1e-9 and dt < 1e-6 and de < 1e-9 and
                  res.iterations == iters and res_tol.iterations == o_tol.iterations and
                  np.array_equal(res_tol.indices.cpu().numpy(), o_tol.indices[-1]))
            ok_all &= ok
            print(f"exchange={exchange} graph={graph} world={world}: state bit-identical across ranks {same}, src64 "
                  f"bit-identical {same_src}, same stop decision {same_its} ({res_tol.iterations} iterations, oracle "
                  f"{o_tol.iterations}); last-iteration indices == oracle {ok_idx}; |dR|={dR:.2e} |dt|={dt:.2e} mm "
                  f"|derr|={de:.2e}: {'OK' if ok else 'FAIL'}", flush=True)
        if s2m.peer is not None:
            s2m.peer.close()
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    if rank == 0:
        print("S2M MULTI-GPU PARITY OK", flush=True)


if __name__ == "__main__":
    main()
