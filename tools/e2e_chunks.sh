#!/bin/bash
# sweep the HostPipeline chunk count (run under gpurun)
for c in 4 8 12 16 24 32; do
  B200ICP_E2E_CHUNKS=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print($c, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
