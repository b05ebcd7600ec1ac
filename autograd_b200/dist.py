"""Batch-sharded gradients across GPUs (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed layer; BASELINE.json's north_star asks for exactly one
thing: batch-separable objectives (minibatch losses such as examples/neural_net.py:45-48 and
examples/convnet.py:61-64) are sharded by batch, every rank runs the unchanged ``grad(objective)`` on its
shard, and the gradient leaves are summed with ONE allreduce over NVLink.  Tape-serial workloads stay on one
GPU.  The collective runs on NCCL tensors that alias our HBM buffers (no staging copy); on the CPU test
suite the same code runs over gloo.
"""

from __future__ import annotations

import os

import numpy as np

from . import backend, lazy
from .array import DeviceArray, concatenate, ravel, reshape

_pg = None


def init(backend_name: str | None = None):
    """Join the torch.distributed job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    global _pg
    import torch
    import torch.distributed as dist

    if dist.is_initialized():
        _pg = dist
        return dist
    be = backend.get()
    if backend_name is None:
        backend_name = "nccl" if be.name == "cuda" else "gloo"
    if backend_name == "nccl":
        torch.cuda.set_device(be.device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", be.device))
    else:
        dist.init_process_group(backend_name)
    _pg = dist
    return dist


def world_size() -> int:
    return int(os.environ.get("WORLD_SIZE", "1")) if _pg is None else _pg.get_world_size()


def rank() -> int:
    return int(os.environ.get("RANK", "0")) if _pg is None else _pg.get_rank()


def _as_torch(x: DeviceArray):
    """A torch tensor aliasing the payload of a contiguous DeviceArray."""
    import torch

    be = backend.get()
    m = x._mat()
    if not m.is_contiguous:
        raise ValueError("allreduce needs a contiguous array")
    if be.name == "cuda":
        return torch.as_tensor(x, device=torch.device("cuda", be.device))
    el, e0 = be._elems(m.ptr, m.dtype)  # host emulation used by the CPU test-suite
    return torch.from_numpy(el[e0: e0 + x.size]).reshape(m.shape)


_peer = {"ready": False, "slot_bytes": 0, "transport": None, "nvls": False, "keep": None}


def _attach_symmetric(be, world: int, r: int, max_bytes: int) -> bool:
    """Comm blocks from torch.distributed's symmetric memory: peer mappings AND, where the platform has NVLS, a
    multicast mapping of the same blocks - the plumbing (handle exchange, cuMulticast* calls) is torch's, the kernels
    that use the addresses are ours (csrc/comm.cu).  False when this torch build / platform cannot provide it."""
    import torch

    try:
        import torch.distributed._symmetric_memory as symm

        nbytes = be.comm_block_bytes(world, max_bytes)
        dev = torch.device("cuda", be.device)
        block = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        block.zero_()
        torch.cuda.synchronize(dev)
        hdl = symm.rendezvous(block, _pg.group.WORLD)
        peers = [int(p) for p in hdl.buffer_ptrs]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if len(peers) != world or not all(peers):
            return False
        _pg.barrier()                                # every rank zeroed its block before anybody raises a flag in it
        be.comm_attach(r, world, max_bytes, peers, mc)
        _peer["keep"] = (block, hdl)                 # the mappings live as long as these objects
        _peer["transport"], _peer["nvls"] = "torch symmetric memory", bool(mc)
        return True
    except Exception as exc:                         # older torch, no fabric / pidfd support in the container, ...
        _peer["transport_error"] = f"{type(exc).__name__}: {exc}"
        return False


def enable_peer_allreduce(max_bytes: int = 4 << 20, symmetric: bool = True) -> bool:
    """Set up the one-kernel NVLink allreduce (csrc/comm.cu) for float64 vectors up to ``max_bytes``.  The comm blocks
    come from torch's symmetric memory when available (which also gives the NVLS multicast mapping used by the
    in-switch reduction); otherwise every rank cudaMallocs its block and the CUDA-IPC handles are exchanged through
    torch.distributed."""
    if _pg is None or _pg.get_world_size() == 1 or backend.get().name != "cuda":
        return False
    if _peer["ready"]:
        return True
    be = backend.get()
    world, r = _pg.get_world_size(), _pg.get_rank()
    ok = symmetric and os.environ.get("B200AG_COMM_SYMMETRIC", "1") != "0" and _attach_symmetric(be, world, r, max_bytes)
    # all ranks must take the same path
    import torch

    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", be.device))
    _pg.all_reduce(flag, op=_pg.ReduceOp.MIN)
    if ok and int(flag.item()) == 0:
        be.comm_destroy()
        _peer["keep"] = None
        ok = False
    if not ok:
        handle = be.comm_create(r, world, max_bytes)
        gathered = [None] * world
        _pg.all_gather_object(gathered, handle)
        be.comm_connect(b"".join(gathered))
        _peer["transport"], _peer["nvls"] = "CUDA IPC", False
    _pg.barrier()
    _peer["ready"], _peer["slot_bytes"] = True, max_bytes
    return True


def allreduce_sum_(x: DeviceArray) -> DeviceArray:
    """In-place sum over ranks of a materialised, contiguous array: the one-kernel peer-memory allreduce when it is
    enabled and the vector fits, otherwise ncclAllReduce on the library stream."""
    if _pg is None or _pg.get_world_size() == 1:
        return x
    import torch

    x._own_buffer()
    x._mat().buf.inexact = True        # a float sum over ranks
    be = backend.get()
    if _peer["ready"] and x.dtype == lazy.F64 and (x.size + 1) * 8 <= _peer["slot_bytes"]:
        m = x._mat()
        if m.is_contiguous:
            be.comm_allreduce_sum_f64(m.ptr, m.ptr, x.size)
            return x
    if be.name == "cuda":
        stream = torch.cuda.ExternalStream(be.stream_handle(), device=torch.device("cuda", be.device))
        with torch.cuda.stream(stream):
            t = _as_torch(x)
            _pg.all_reduce(t, op=_pg.ReduceOp.SUM)
    else:
        _pg.all_reduce(_as_torch(x), op=_pg.ReduceOp.SUM)
    return x


def _leaves(tree, acc):
    if isinstance(tree, DeviceArray):
        acc.append(tree)
    elif isinstance(tree, (list, tuple)):
        for t in tree:
            _leaves(t, acc)
    elif isinstance(tree, dict):
        for k in sorted(tree):
            _leaves(tree[k], acc)
    else:
        raise TypeError(f"gradient leaf of type {type(tree).__name__} is not a DeviceArray")


def _rebuild(tree, it):
    if isinstance(tree, DeviceArray):
        return next(it)
    if isinstance(tree, (list, tuple)):
        return type(tree)(_rebuild(t, it) for t in tree)
    return {k: _rebuild(tree[k], it) for k in sorted(tree)}


def flatten_tree(tree) -> DeviceArray:
    """All gradient leaves raveled into ONE contiguous vector (one message instead of one per leaf)."""
    acc = []
    _leaves(tree, acc)
    dt = np.result_type(*[a.dtype for a in acc])
    return concatenate([ravel(a) if a.dtype == dt else ravel(a).astype(dt) for a in acc])


def unflatten_like(flat: DeviceArray, tree):
    acc = []
    _leaves(tree, acc)
    out, pos = [], 0
    for a in acc:
        piece = reshape(flat[pos: pos + a.size], a.shape)
        out.append(piece if piece.dtype == a.dtype else piece.astype(a.dtype))
        pos += a.size
    return _rebuild(tree, iter(out))


def allreduce_grads(grads):
    """Sum a gradient tree over ranks with a single collective; returns a tree of the same structure."""
    if _pg is None or _pg.get_world_size() == 1:
        return grads
    flat = flatten_tree(grads)
    allreduce_sum_(flat)
    return unflatten_like(flat, grads)


def shard_batch(n_total: int):
    """This rank's contiguous slice of a batch of ``n_total`` examples (even split, remainder to low ranks)."""
    w, r = world_size(), rank()
    base, extra = divmod(n_total, w)
    lo = r * base + min(r, extra)
    return slice(lo, lo + base + (1 if r < extra else 0))
