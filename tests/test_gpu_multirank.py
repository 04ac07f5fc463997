"""GPU, world size >= 2 over NCCL: the three shardings of the path against the single-GPU evaluation, bit for bit
(VERDICT r1 missing 4).  Spawns one process per GPU with torch.distributed.run; skipped on a one-GPU box
(the world-size-2 gloo tests in test_host_logic.py and the one-rank NCCL test in test_gpu_lml.py cover the host logic there).
Run on two GPUs of one box:  gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_multirank.py -m gpu -q'"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_shardings_are_bit_identical_to_one_gpu(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() + world) % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    out = os.path.join(os.path.dirname(HERE), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    if r.returncode != 0 or not lines:
        with open(os.path.join(out, f"r02_multirank_{world}gpu_fail.txt"), "w") as f:
            f.write(r.stdout[-6000:] + "\n---- stderr ----\n" + r.stderr[-12000:])
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-4000:])
    rep = json.loads(lines[-1])
    assert rep["all_ranks_ok"] and rep["world"] == world and "failed" not in rep, rep
    with open(os.path.join(out, f"r02_multirank_{world}gpu.json"), "w") as f:
        json.dump(rep, f, indent=1)
