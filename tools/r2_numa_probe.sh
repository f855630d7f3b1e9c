#!/bin/bash
# Host topology of the GPU box + A/B of the end-to-end bench with and without NUMA-local pinned tables.
# usage (8-GPU box): bash tools/r2_numa_probe.sh 8 > gpurun_out/r2_numa_probe.txt 2>&1
N=${1:-8}
echo "== lscpu"; lscpu | grep -i -E "model name|socket|numa|^cpu\(s\)|thread"
echo "== nodes online: $(cat /sys/devices/system/node/online 2>/dev/null)"
for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== GPU PCI functions and their NUMA node"
for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do
  a=$(echo ${b#0000} | tr 'A-Z' 'a-z'); echo "$b numa_node=$(cat /sys/bus/pci/devices/$a/numa_node 2>/dev/null)"; done
echo "== nvidia-smi topo -m"; nvidia-smi topo -m
echo "== affinity of this shell: $(taskset -p $$ 2>/dev/null)"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 \
          bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --no-cpu-baseline "${@:2}" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); e = d['e2e']
        print(json.dumps({k: e[k] for k in ('value', 'ms_per_step', 'h2d_only_ms', 'h2d_only_gbs_per_gpu', 'host_numa')}), 'resident ms', d['ms_per_step'])
"; }
echo "== e2e, pinned tables wherever the process runs"; run 29541 --no-numa-bind
echo "== e2e, pinned tables on the GPU's NUMA node (default)"; run 29542
echo "== same, 4 chunks"; run 29543 --e2e-chunks 4
