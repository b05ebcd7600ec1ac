"""N>1 path on CPU: world_size-2 gloo job, batch-sharded gradients + one allreduce (no GPU needed)."""

import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_data_parallel_gradients_match_full_batch_world2():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
        assert f"rank {rank}/2 ok" in out


def test_shard_batch_covers_everything(monkeypatch):
    from autograd_b200 import dist

    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            monkeypatch.setenv("WORLD_SIZE", str(world))
            monkeypatch.setenv("RANK", str(r))
            sl = dist.shard_batch(65)
            seen.extend(range(65)[sl])
        assert seen == list(range(65))
