"""Multi-GPU tests (need >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests -m gpu`); skipped on one GPU."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("world", [2])
def test_mcu_row_shards_over_nccl(world):
    """BASELINE config 5 at test size: one process per GPU, the four small exchanges over NCCL."""
    _need(world)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_port()),
                        os.path.join(ROOT, "tests", "_mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": OK") == 4, r.stdout


def test_batch_round_robin_two_ranks():
    """BASELINE config 4 at test size through bench.py's own multi-rank path (round-robin, no collective
    on the data path): the run must verify image 0 against the oracle and print one JSON line."""
    _need(2)
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--images", "48", "--steps", "2",
                        "--warmup", "1", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["config"]["images_per_rank"] == 24
    assert line["config"]["verified"] == "image 0 byte-identical to oracle"
    assert line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
