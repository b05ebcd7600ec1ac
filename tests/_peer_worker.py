"""Worker for the multi-GPU peer-memory allreduce check (run under torchrun, one rank per GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as onp

import autograd_b200 as b200

b200.install()
from autograd_b200 import dist

dist.init("nccl")
world, rank = dist.world_size(), dist.rank()
be = b200.backend.get()
assert dist.enable_peer_allreduce(4 << 20)

rs = onp.random.RandomState(1234)
all_vecs = [rs.randn(178110) for _ in range(world)]          # identical on every rank (same seed)
want = onp.zeros(178110)
for v in all_vecs:                                            # rank order, like the kernel
    want = want + v

# 1. eager calls, many epochs (exercises both slots and the epoch counter), one-shot and two-shot interleaved,
#    plus an odd length and a short vector through both algorithms
import torch

NVLS = bool(dist._peer["nvls"])
if rank == 0:
    print(f"comm blocks: {dist._peer['transport']}, NVLS multicast mapping: {NVLS}"
          + (f" (symmetric memory unavailable: {dist._peer.get('transport_error')})" if dist._peer.get("transport_error") else ""),
          flush=True)
for mode in (1, 2, 3) + ((4,) if NVLS else ()):
    be.comm_mode(mode)
    for n_el in (178109, 7, 2):
        xs = b200.asarray(all_vecs[rank][:n_el].copy())
        dist.allreduce_sum_(xs)
        exp = onp.zeros(n_el)
        for v in all_vecs:
            exp = exp + v[:n_el]
        got = xs.get()
        if mode == 4:
            # the NVSwitch adds in its own order: equal to rounding, and the SAME bits on every rank
            assert onp.allclose(got, exp, rtol=1e-14, atol=1e-14), f"rank {rank} NVLS n {n_el}: {onp.abs(got - exp).max()}"
            t = torch.as_tensor(xs, device="cuda")
            gathered = [torch.empty_like(t) for _ in range(world)]
            dist._pg.all_gather(gathered, t)
            assert all(torch.equal(g, gathered[0]) for g in gathered), f"NVLS result differs between ranks (n {n_el})"
        else:
            assert onp.array_equal(got, exp), f"rank {rank} mode {mode} n {n_el}"
be.comm_mode(0)
for it in range(20):
    be.comm_mode(1 + (it % 3))
    x = b200.asarray(all_vecs[rank] * (it + 1))
    dist.allreduce_sum_(x)
    got = x.get()
    exp = onp.zeros(178110)
    for v in all_vecs:
        exp = exp + v * (it + 1)                              # same order, same roundings as the kernel
    if not onp.array_equal(got, exp):
        # diagnose: which multiple of each rank's vector did we actually sum?
        A = onp.stack(all_vecs, axis=1)
        coef = onp.linalg.lstsq(A[:2000], got[:2000], rcond=None)[0]
        raise AssertionError(f"rank {rank} iter {it}: summed coefficients {coef} (expected {it + 1} each), status {be.comm_status()}")
assert be.comm_status() == 0
be.comm_mode(0)

# 2. bit-identical across ranks
t = torch.as_tensor(x, device="cuda")
gathered = [torch.empty_like(t) for _ in range(world)]
dist._pg.all_gather(gathered, t)
for g in gathered:
    assert torch.equal(g, gathered[0])

# 3. inside a replayed CUDA graph: compute + allreduce in ONE graph launch
xin = b200.asarray(all_vecs[rank])


def step(v):
    y = v * 2.0 + 1.0
    y = y.materialize()
    dist.allreduce_sum_(y)
    return y


fast = b200.capture(step)
for it in range(6):
    out = fast(xin).get()
    ref = want * 2.0 + world
    assert onp.allclose(out, ref, rtol=1e-14, atol=1e-12), f"graph replay {it}: {onp.abs(out - ref).max()}"
assert be.comm_status() == 0

# 4. latency vs NCCL
def timeit(fn, reps=200):
    for _ in range(20):
        fn()
    be.synchronize(); dist._pg.barrier()
    e0, e1 = be.event_create(), be.event_create()
    be.event_record(e0)
    for _ in range(reps):
        fn()
    be.event_record(e1); be.synchronize()
    return be.event_elapsed_ms(e0, e1) / reps * 1e3

y = b200.asarray(all_vecs[rank]).materialize()
be.comm_mode(1)
t_one = timeit(lambda: dist.allreduce_sum_(y))
be.comm_mode(2)
t_two = timeit(lambda: dist.allreduce_sum_(y))
be.comm_mode(3)
t_push = timeit(lambda: dist.allreduce_sum_(y))
t_nvls = None
if NVLS:
    be.comm_mode(4)
    t_nvls = timeit(lambda: dist.allreduce_sum_(y))
    # NVLS inside a replayed graph
    fast4 = b200.capture(step)
    for it in range(4):
        out = fast4(xin).get()
        assert onp.allclose(out, want * 2.0 + world, rtol=1e-14, atol=1e-12), f"NVLS graph replay {it}"
be.comm_mode(0)
t_peer = timeit(lambda: dist.allreduce_sum_(y))
dist._peer["ready"] = False
t_nccl = timeit(lambda: dist.allreduce_sum_(y))
dist._peer["ready"] = True
# latency against message size (pull / NVLS / NCCL): what a split of the gradient into an early and a late message could hide
sizes = {}
for n_el in (2048, 22000, 178110):
    yy = b200.asarray(all_vecs[rank][:n_el].copy()).materialize()
    row = {}
    for name, mode in (("pull", 1), ("push", 3)) + ((("nvls", 4),) if NVLS else ()):
        be.comm_mode(mode)
        row[name] = timeit(lambda: dist.allreduce_sum_(yy), reps=100)
    be.comm_mode(0)
    sizes[n_el] = row
if rank == 0:
    print("allreduce latency by size (us):", {k: {a: round(b, 1) for a, b in v.items()} for k, v in sizes.items()}, flush=True)
    print(f"peer-memory allreduce 1.42 MB x{world}: one-shot pull {t_one:.1f} us, two-shot {t_two:.1f} us, push {t_push:.1f} us, automatic {t_peer:.1f} us/call; "
          f"NVLS multimem {'%.1f us' % t_nvls if t_nvls is not None else 'n/a'}; NCCL via torch: {t_nccl:.1f} us/call", flush=True)
print(f"rank {rank}/{world} peer allreduce ok", flush=True)
dist._pg.barrier()
