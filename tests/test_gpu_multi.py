"""The one data-plane exchange of the path on real GPUs (north_star: "NCCL only gathers per-GPU point counts and merged clouds
for the fusion config"): two ranks under torchrun, NCCL backend.  Needs two visible GPUs (`gpurun --gpus 2`); the gloo
world-size-2 test in test_host_cpu.py covers the same host logic on CPU."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_fused_clouds_gather_to_rank0_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi", "nccl_gather_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["ok"] and rec["world"] == 2 and rec["backend"] == "nccl", rec
    assert all(s > 1000 for s in rec["sizes"]), rec
