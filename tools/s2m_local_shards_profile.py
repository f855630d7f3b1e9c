"""Emulate the 8-rank scan-to-map protocol on ONE GPU (8 local shards of the 2^24-point map) so the
per-rank kernel times can be read from an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/s2m_local_shards_profile.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_slam_yolo_b200 as m
from oracle import icp_oracle as orc

M, N, G = 1 << 24, 8192, 8
full = orc.synth_map(M)
shards = []
for g in range(G):
    b, e = m.shard_range(M, g, G)
    shards.append(m.MapShard(torch.from_numpy(full[b:e]).cuda(), global_offset=b))
run = m.scan_to_map.ScanToMapLocalShards(shards, N)
run.init(torch.from_numpy(orc.synth_scan_for_map(N)).cuda())
for it in range(int(os.environ.get("ITERS", "4"))):
    run.step(30, -1.0)
torch.cuda.synchronize()
print("error after", run.result().iterations, "iterations:", run.result().error)
