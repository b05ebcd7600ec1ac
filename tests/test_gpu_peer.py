"""Multi-GPU collective under `pytest -m gpu`: the peer-memory / NVLS allreduce of csrc/comm.cu on 2 ranks (one
process per GPU over torch.distributed), skipped on boxes with a single GPU.  tests/_peer_worker.py checks every
algorithm against the rank-ordered host sum (bit for bit for the peer kernels, to rounding + cross-rank bit-identity
for the in-switch reduction), many epochs, and a CUDA graph that contains the collective."""

import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_peer_allreduce_world2():
    if _gpu_count() < 2:
        pytest.skip("needs at least 2 GPUs on the box")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        env.pop("B200AG_DEVICE", None)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_peer_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=420)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-4000:]}"
        assert f"rank {rank}/2 peer allreduce ok" in out
