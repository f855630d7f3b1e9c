"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): the scan-to-map path with the
map sharded over the ranks, run as a torchrun job (tools/check_s2m_multigpu.py): state
bit-identical across ranks, indices / pose / error equal to the CPU oracle, for the peer-store
exchange, its CUDA-graph replay and the NCCL exchange.  The job's report is kept under
gpurun_out/ so that it can be archived in profiles/."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("world", [2, 8])
def test_scan_to_map_sharded_over_ranks_matches_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world),
           os.path.join(ROOT, "tools", "check_s2m_multigpu.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"s2m_multigpu_parity_{world}gpu.txt"), "w") as f:
        f.write(proc.stdout + "\n--- stderr ---\n" + proc.stderr[-4000:])
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "S2M MULTI-GPU PARITY OK" in proc.stdout
    assert proc.stdout.count(": OK") == 3
