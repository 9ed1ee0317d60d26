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
    for attempt in range(3):   # the free port found by _port() can be taken again before torchrun binds it
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(_port()),
                            os.path.join(ROOT, "tests", "_mgpu_worker.py")], capture_output=True, text=True, timeout=600)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            break
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": OK") == 4, r.stdout


def test_batch_round_robin_two_ranks():
    """BASELINE config 4 at test size through bench.py's own multi-rank path (round-robin, no collective
    on the data path): the run must verify image 0 against the oracle and print one JSON line."""
    _need(2)
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--images", "48", "--steps", "2",
                        "--warmup", "1", "--no-cpu-baseline", "--config5-size", "4096"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["config"]["images_per_rank"] == 24
    assert "byte-identical to the oracle" in line["config"]["verified"]
    assert line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
    assert 0.2 < line["e2e"]["frac_of_copy_ceiling"] < 1.3
    # BASELINE config 5 through the same run: sharded over the two ranks, checked against the ORACLE's digest
    c5 = line["extra"]["config5"]
    assert c5["n_gpus"] == 2 and "equal the oracle's file" in c5["verified"], c5


def test_c_abi_sharded_over_two_devices():
    """dmmt_encode_sharded with one context per DEVICE of this process: the exchanges are peer loads between the two
    GPUs, K4 of the second shard stores into the file in the first GPU's memory over NVLink."""
    _need(2)
    import numpy as np

    import dmmt_jpeg_encoder_b200 as D
    from conftest import synth_image
    from oracle import oracle as O

    c0, c1 = D.Context(0), D.Context(1)
    try:
        for (w, h, preset) in [(1000, 650, 2), (333, 97, 0), (640, 360, 1)]:
            px = synth_image("photo", w, h, 3)
            want = O.encode(px, 255, preset).jpeg
            assert c0.encode_sharded(px, 2, 255, D.Options(preset, 8, 0), contexts=[c0, c1]) == want
            assert c0.encode_sharded(px, 4, 255, D.Options(preset, 8, 0), contexts=[c0, c1, c1, c0]) == want
        noise = np.random.default_rng(1).integers(0, 256, (160, 272, 3), dtype=np.uint8)      # overflow -> retry
        assert c0.encode_sharded(noise, 2, 255, D.Options(2, 8, 1), contexts=[c0, c1]) == O.encode(noise, 255, 2, 8, 1).jpeg
    finally:
        c0.close()
        c1.close()
