"""Secondary measurements reported under ``other_configs`` by ``bench.py --extras``:
the remaining BASELINE.json configs (C1 MLP, C3 LSTM, C4 fluidsim, C5 convnet) at their named sizes (or the
largest size that fits, stated), each with the reference's NumPy path timed beside it on a bounded sample,
plus achieved bandwidth / FLOP rate of representative kernels measured with CUDA events on the library stream.
"""

from __future__ import annotations

import os
import time

import numpy as onp

import autograd_b200 as b200
from autograd import grad, value_and_grad

from examples import workloads as w

HBM_PEAK = None


def _peak():
    global HBM_PEAK
    if HBM_PEAK is None:
        import json

        try:
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
                HBM_PEAK = float(json.load(f)["hbm_gbs"])
        except Exception:
            HBM_PEAK = 6650.0
    return HBM_PEAK


def tree_to_dev(t):
    if isinstance(t, onp.ndarray):
        return b200.asarray(t)
    if isinstance(t, (list, tuple)):
        return type(t)(tree_to_dev(e) for e in t)
    if isinstance(t, dict):
        return {k: tree_to_dev(v) for k, v in t.items()}
    return t


def tree_leaves(t):
    if isinstance(t, (list, tuple)):
        return [l for x in t for l in tree_leaves(x)]
    if isinstance(t, dict):
        return [l for k in sorted(t) for l in tree_leaves(t[k])]
    return [t]


def materialize(t):
    for leaf in tree_leaves(t):
        if isinstance(leaf, b200.DeviceArray):
            leaf.materialize()
    return t


def time_device(be, fn, warmup=2, reps=5):
    """ms per call (CUDA events on the library stream, sync on both sides), launches per call, peak memory."""
    for _ in range(warmup):
        materialize(fn())
    be.synchronize()
    be.mem_stats()
    e0, e1 = be.event_create(), be.event_create()
    l0 = be.launch_count()
    t0 = time.perf_counter()
    be.event_record(e0)
    for _ in range(reps):
        materialize(fn())
    be.event_record(e1)
    be.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms = be.event_elapsed_ms(e0, e1) / reps
    return {"ms": ms, "wall_ms": wall, "launches": (be.launch_count() - l0) // reps,
            "peak_mem_gb": be.mem_stats()["peak_in_use"] / 1e9}


def time_replay(be, fn, args, reps=10):
    """The same call replayed as a CUDA graph (autograd_b200.capture): ms per call and graph size."""
    try:
        fast = b200.capture(fn)
        fast(*args)  # warm-up + capture + first replay
        be.synchronize()
        e0, e1 = be.event_create(), be.event_create()
        t0 = time.perf_counter()
        be.event_record(e0)
        for _ in range(reps):
            fast(*args)
        be.event_record(e1)
        be.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms = be.event_elapsed_ms(e0, e1) / reps
        nodes = list(fast.num_nodes().values())[0]
        out = fast(*args)
        return {"ms": ms, "wall_ms": wall, "graph_nodes": nodes, "grad_evals_per_sec": 1e3 / max(ms, wall)}, out
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}, None


def time_host(fn, warmup=1, reps=2):
    with b200.host_arrays(), onp.errstate(all="ignore"):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
    return min(ts)


def max_rel_err(got, want):
    worst = 0.0
    for g, r in zip(tree_leaves(got), tree_leaves(want)):
        g = g.get() if isinstance(g, b200.DeviceArray) else onp.asarray(g)
        r = onp.asarray(r)
        worst = max(worst, float(onp.max(onp.abs(g.astype(onp.float64) - r.astype(onp.float64))) /
                               max(float(onp.max(onp.abs(r))), 1e-300)))
    return worst


# ----------------------------------------------------------------------------- configs
def c1_mlp(be):
    params, X, T = w.mlp_data()
    dp, dX, dT = tree_to_dev(params), b200.asarray(X), b200.asarray(T)
    g = grad(w.mlp_objective)
    dev = time_device(be, lambda: g(dp, dX, dT, 1.0), warmup=3, reps=20)
    cpu = time_host(lambda: g(params, X, T, 1.0), warmup=2, reps=10)
    with b200.host_arrays():
        want = g(params, X, T, 1.0)
    err = max_rel_err(g(dp, dX, dT, 1.0), want)
    replay, rout = time_replay(be, g, (dp, dX, dT, 1.0), reps=50)
    if rout is not None:
        replay["max_rel_err_vs_reference"] = max_rel_err(rout, want)
    return {"workload": "C1 MLP [784,200,100,10] batch 256 f64, grad of -log_posterior", "grad_evals_per_sec": 1e3 / dev["ms"],
            **dev, "graph_replay": replay, "flops_per_grad": 193e6, "gflops": 193e6 / (dev["ms"] * 1e-3) / 1e9,
            "cpu_reference": {"grad_evals_per_sec": 1 / cpu, "seconds": cpu, "sample": "full size", "cores_blas": os.cpu_count()},
            "max_rel_err_vs_reference": err}


def c3_lstm(be, T=256, B=1024, H=512, C=128, cpu_T=2):
    params, X, Y = w.lstm_data(seq_len=T, batch=B, hidden=H, chars=C)
    dp, dX = tree_to_dev(params), b200.asarray(X)
    g = grad(w.lstm_objective)
    dev = time_device(be, lambda: g(dp, dX, dX), warmup=1, reps=2)
    replay, _ = time_replay(be, g, (dp, dX, dX), reps=3)
    pc, Xc, _ = w.lstm_data(seq_len=cpu_T, batch=B, hidden=H, chars=C)
    cpu = time_host(lambda: g(pc, Xc, Xc), warmup=0, reps=1)
    cpu_full = cpu * T / cpu_T
    with b200.host_arrays():
        want = g(pc, Xc, Xc)
    err = max_rel_err(g(tree_to_dev(pc), b200.asarray(Xc), b200.asarray(Xc)), want)
    return {"workload": f"C3 LSTM BPTT hidden {H} seq {T} batch {B} f32 params (f64 activations by NumPy promotion)",
            "grad_evals_per_sec": 1e3 / dev["ms"], **dev, "graph_replay": replay, "flops_per_grad": 2.17e12 * T / 256,
            "tflops_fp64": 2.17e12 * T / 256 / (dev["ms"] * 1e-3) / 1e12,
            "cpu_reference": {"grad_evals_per_sec": 1 / cpu_full, "seconds_extrapolated": cpu_full,
                              "sample": f"T={cpu_T} of {T} time steps at full width ({cpu:.2f} s), scaled linearly in T",
                              "cores_blas": os.cpu_count()},
            "max_rel_err_vs_reference_at_sample": err}


def c4_fluid(be, rows=2048, cols=2048, steps=200, cpu_rows=256, cpu_steps=2, replay_ok=1):
    p, s0, tg = w.fluid_data(rows, cols)
    dp, ds, dt = b200.asarray(p), b200.asarray(s0), b200.asarray(tg)
    vg = value_and_grad(w.fluid_objective)
    dev = time_device(be, lambda: vg(dp, ds, dt, rows, cols, steps), warmup=1, reps=1)
    replay, _ = time_replay(be, vg, (dp, ds, dt, rows, cols, steps), reps=2) if replay_ok else ({"skipped": "memory"}, None)
    pc, sc, tc = w.fluid_data(cpu_rows, cpu_rows)
    cpu = time_host(lambda: vg(pc, sc, tc, cpu_rows, cpu_rows, cpu_steps), warmup=0, reps=1)
    cpu_full = cpu * (rows * cols * steps) / (cpu_rows * cpu_rows * cpu_steps)
    with b200.host_arrays():
        want = vg(pc, sc, tc, cpu_rows, cpu_rows, cpu_steps)
    got = vg(b200.asarray(pc), b200.asarray(sc), b200.asarray(tc), cpu_rows, cpu_rows, cpu_steps)
    err = max_rel_err(got[1], want[1])
    return {"workload": f"C4 fluidsim value_and_grad through {steps} steps of a {rows}x{cols} f64 grid",
            "grad_evals_per_sec": 1e3 / dev["ms"], **dev, "graph_replay": replay,
            "cell_steps_per_sec": rows * cols * steps / (dev["ms"] * 1e-3),
            "cpu_reference": {"grad_evals_per_sec": 1 / cpu_full, "seconds_extrapolated": cpu_full,
                              "sample": f"{cpu_rows}x{cpu_rows} grid x {cpu_steps} steps ({cpu:.2f} s), scaled by cells x steps",
                              "cores": 1},
            "max_rel_err_vs_reference_at_sample": err}


def c5_convnet(be, N=65536, cpu_N=64):
    n, _pred, loss = w.convnet_make()
    W, X, T = w.convnet_data(N, n)
    dW, dX, dT = b200.asarray(W), b200.asarray(X), b200.asarray(T)
    g = grad(loss)
    dev = time_device(be, lambda: g(dW, dX, dT), warmup=1, reps=3)
    replay, _ = time_replay(be, g, (dW, dX, dT), reps=5)
    Wc, Xc, Tc = w.convnet_data(cpu_N, n)
    cpu = time_host(lambda: g(Wc, Xc, Tc), warmup=0, reps=1)
    cpu_full = cpu * N / cpu_N
    with b200.host_arrays():
        want = g(Wc, Xc, Tc)
    err = max_rel_err(g(b200.asarray(Wc), b200.asarray(Xc), b200.asarray(Tc)), want)
    return {"workload": f"C5 LeNet convnet grad, batch {N}, 28x28 f64", "grad_evals_per_sec": 1e3 / dev["ms"], **dev,
            "graph_replay": replay,
            "images_per_sec": N / (dev["ms"] * 1e-3), "gflops": 1.5e6 * N / (dev["ms"] * 1e-3) / 1e9,
            "cpu_reference": {"grad_evals_per_sec": 1 / cpu_full, "seconds_extrapolated": cpu_full,
                              "sample": f"batch {cpu_N} ({cpu:.2f} s), scaled linearly in batch", "cores": 1},
            "max_rel_err_vs_reference_at_sample": err}


# ----------------------------------------------------------------------------- kernel micro-measurements
def kernels(be):
    """Achieved GB/s (algorithmic bytes / CUDA-event time) of representative hot-path kernels, inputs > L2."""
    import autograd.numpy as np
    from autograd import elementwise_grad as egrad

    out = {}
    n = 1 << 27  # 1 GiB per f64 vector
    x = b200.asarray(onp.linspace(-3.0, 3.0, n))
    y = b200.asarray(onp.linspace(1.0, 2.0, n))

    def rate(fn, nbytes, reps=10):
        r = time_device(be, fn, warmup=3, reps=reps)
        gbs = nbytes / (r["ms"] * 1e-3) / 1e9
        return {"ms": r["ms"], "GB/s": gbs, "frac_of_hbm_peak": gbs / _peak(), "launches": r["launches"]}

    out["vjp_multiply (y*g): 2 reads + 1 write"] = rate(lambda: y * x, 24 * n)
    out["vjp_tanh (g/cosh(x)**2): 2 reads + 1 write"] = rate(lambda: y / np.cosh(x) ** 2, 24 * n)
    out["egrad(tanh) order 1 fused fwd+bwd: 1 read + 1 write"] = rate(lambda: egrad(w.tanh_fn)(x), 16 * n)
    z = b200.asarray(onp.linspace(0.0, 1.0, n))
    out["accumulate (a + b + c): 3 reads + 1 write"] = rate(lambda: x + y + z, 32 * n)
    xm = x.materialize()
    out["sum all (1 read)"] = rate(lambda: np.sum(xm), 8 * n)
    m2 = b200.asarray(onp.random.RandomState(0).rand(8192, 8192))
    nn = 8192 * 8192
    out["column sum [8192,8192] axis 0"] = rate(lambda: np.sum(m2, axis=0), 8 * nn)
    out["row sum [8192,8192] axis 1"] = rate(lambda: np.sum(m2, axis=1), 8 * nn)
    out["max [8192,8192] axis 1"] = rate(lambda: np.max(m2, axis=1), 8 * nn)
    out["transpose copy [8192,8192]"] = rate(lambda: m2.T.copy(), 16 * nn)
    out["roll stencil (4 shifted reads of one array + 1 write)"] = rate(
        lambda: (np.roll(m2, 1, 0) + np.roll(m2, -1, 0) + np.roll(m2, 1, 1) + np.roll(m2, -1, 1)) / 4.0, 16 * nn)
    ri = b200.asarray(onp.random.RandomState(1).randint(0, 8192, nn))
    ci = b200.asarray(onp.random.RandomState(2).randint(0, 8192, nn))
    out["gather f[i,j] random (2 idx + value read + write)"] = rate(lambda: m2[ri, ci], 32 * nn, reps=5)
    g = b200.asarray(onp.random.RandomState(3).rand(nn))

    def scatter():
        z = b200.array.zeros((8192, 8192))
        onp.add.at(z, (ri, ci), g)
        return z

    out["scatter-add np.add.at random (2 idx + value read + 16 B RMW)"] = rate(scatter, 40 * nn, reps=5)
    # FP64 GEMM on the DMMA pipe
    for (M, N, K) in ((4096, 4096, 4096), (1024, 512, 641), (641, 512, 1024), (256, 200, 784)):
        A = b200.asarray(onp.random.RandomState(4).rand(M, K))
        B = b200.asarray(onp.random.RandomState(5).rand(K, N))
        r = time_device(be, lambda: np.dot(A, B), warmup=2, reps=10)
        out[f"dgemm {M}x{N}x{K}"] = {"ms": r["ms"], "TFLOP/s": 2.0 * M * N * K / (r["ms"] * 1e-3) / 1e12}
    # float32 products on tcgen05 (3xTF32): forward shape, and the two VJP shapes G B^T / A^T G (transposed views)
    for (M, N, K) in ((8192, 8192, 4096), (4096, 4096, 4096), (1024, 512, 16384)):
        A = b200.asarray(onp.random.RandomState(4).rand(M, K).astype(onp.float32))
        B = b200.asarray(onp.random.RandomState(5).rand(K, N).astype(onp.float32))
        G = b200.asarray(onp.random.RandomState(6).rand(M, N).astype(onp.float32))
        for name, fn, flops in ((f"sgemm(3xTF32, tcgen05) {M}x{N}x{K}", lambda: np.dot(A, B), 2.0 * M * N * K),
                                (f"  its VJP wrt A (G B^T) {M}x{K}x{N}", lambda: np.dot(G, B.T), 2.0 * M * N * K),
                                (f"  its VJP wrt B (A^T G) {K}x{N}x{M}", lambda: np.dot(A.T, G), 2.0 * M * N * K)):
            r = time_device(be, fn, warmup=2, reps=5)
            tf = flops / (r["ms"] * 1e-3) / 1e12
            out[name] = {"ms": r["ms"], "TFLOP/s fp32-equivalent": tf, "TF32 TFLOP/s": 3 * tf,
                         "frac_of_nominal_tf32_peak": 3 * tf / 1100.0, "launches": r["launches"]}
    return out


# ----------------------------------------------------------------------------- batch-sharded configs (N GPUs)
def _dp_time(be, step, reps, world):
    from autograd_b200 import dist as bdist

    for _ in range(3):
        step()
    be.synchronize()
    if world > 1:
        bdist._pg.barrier()
    e0, e1 = be.event_create(), be.event_create()
    l0 = be.launch_count()
    be.event_record(e0)
    for _ in range(reps):
        step()
    be.event_record(e1)
    be.synchronize()
    ms = be.event_elapsed_ms(e0, e1) / reps
    launches = (be.launch_count() - l0) // reps
    if world > 1:
        import torch

        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        bdist._pg.all_reduce(t, op=bdist._pg.ReduceOp.MAX)
        ms = float(t.item())
    return ms, launches


def c1_dp(be, world=1, rank=0):
    """C1 sharded by batch, WEAK scaling: every rank owns a 256-row minibatch (seed = rank), prior split 1/world,
    one allreduce of the flattened 178,110-element gradient per step."""
    from autograd_b200 import dist as bdist

    params, _, _ = w.mlp_data(seed=0)
    _, X, T = w.mlp_data(seed=rank)
    dp, dX, dT = tree_to_dev(params), b200.asarray(X), b200.asarray(T)
    g = grad(w.mlp_objective)
    local = b200.capture(lambda p, x, t: bdist.flatten_tree(g(p, x, t, 1.0 / world)))

    def step_nccl():
        flat = local(dp, dX, dT)
        bdist.allreduce_sum_(flat)
        return flat

    peer = world > 1 and bdist.enable_peer_allreduce(4 << 20)
    variants = {}
    step = step_nccl
    if peer:
        # variant A: forward + backward + flatten replayed as a graph, then ncclAllReduce (NVLS in-switch reduction at 8 ranks)
        bdist._peer["ready"] = False
        variants["graph + ncclAllReduce"] = _dp_time(be, step_nccl, 50, world)
        bdist._peer["ready"] = True

        # variant B: forward + backward + flatten + the one-kernel NVLink peer-memory allreduce in ONE CUDA graph
        def fused(p, x, t):
            flat = bdist.flatten_tree(g(p, x, t, 1.0 / world)).materialize()
            return bdist.allreduce_sum_(flat)

        steps = {}
        be.comm_mode(1 if bdist._peer["nvls"] else 0)
        whole = b200.capture(fused)
        steps["one graph incl. peer-memory allreduce kernel"] = lambda: whole(dp, dX, dT)
        variants["one graph incl. peer-memory allreduce kernel"] = _dp_time(be, steps["one graph incl. peer-memory allreduce kernel"], 50, world)
        if bdist._peer["nvls"]:
            # variant C: the same single graph, the sum taken inside the NVSwitch (multimem.ld_reduce / multimem.st)
            be.comm_mode(4)
            whole_nvls = b200.capture(fused)
            name = "one graph incl. NVLS in-switch allreduce kernel"
            steps[name] = lambda: whole_nvls(dp, dX, dT)
            variants[name] = _dp_time(be, steps[name], 50, world)
        be.comm_mode(0)
        best = min(variants, key=lambda k: variants[k][0])
        if best.startswith("graph + nccl"):
            bdist._peer["ready"] = False
        else:
            step = steps[best]
        ms, launches = variants[best]
    else:
        best = "ncclAllReduce after the replayed graph" if world > 1 else "none (1 GPU)"
        ms, launches = _dp_time(be, step, 50, world)
    res = {"workload": f"C1 MLP grad, 256 rows per GPU x {world} GPUs, 1.42 MB f64 gradient allreduce per step",
           "collective": best, "scaling": "weak", "ms_per_step": ms, "grad_evals_per_sec": world * 1e3 / ms,
           "rows_per_sec": world * 256 * 1e3 / ms, "launches_per_step": launches,
           "variants_ms_per_step": {k: v[0] for k, v in variants.items()},
           "comm_blocks": bdist._peer.get("transport"), "nvls_multicast": bool(bdist._peer.get("nvls"))}
    # parity: the summed gradient equals the reference gradient of the concatenated batch
    flat = step().get()
    if rank == 0:
        Xs, Ts = zip(*[w.mlp_data(seed=r)[1:] for r in range(world)])
        with b200.host_arrays():
            want = g(params, onp.concatenate(Xs), onp.concatenate(Ts), 1.0)
        wflat = onp.concatenate([a.ravel() for wb in want for a in wb])
        res["max_rel_err_vs_reference_full_batch"] = float(onp.max(onp.abs(flat - wflat)) / onp.max(onp.abs(wflat)))
    bdist._peer["ready"] = bool(peer)
    return res


def c5_dp(be, world=1, rank=0, N=65536):
    """C5 sharded by batch, STRONG scaling: N images split evenly over the ranks, one allreduce of the
    44,426-element gradient."""
    from autograd_b200 import dist as bdist

    n, _pred, loss = w.convnet_make()
    per = N // world
    W, _, _ = w.convnet_data(1, n, seed=0)
    _, X, T = w.convnet_data(per, n, seed=100 + rank)
    dW, dX, dT = b200.asarray(W), b200.asarray(X), b200.asarray(T)
    g = grad(loss)
    peer = world > 1 and bdist.enable_peer_allreduce(4 << 20)

    def fused(wv, x, t):
        flat = g(wv, x, t, 1.0 / world).materialize()
        return bdist.allreduce_sum_(flat) if peer else flat

    local = b200.capture(fused)

    def step():
        flat = local(dW, dX, dT)
        if world > 1 and not peer:
            bdist.allreduce_sum_(flat)
        return flat

    ms, launches = _dp_time(be, step, 5, world)
    # parity of the sharded path at a size the host can check: 64 images per rank through the same step function,
    # against the reference gradient of the concatenated batch (the example's logsumexp shifts by the shard's max
    # instead of the global one, convnet.py:40-42: rounding-level differences only)
    _, Xs, Ts = w.convnet_data(64, n, seed=100 + rank)
    got = fused(dW, b200.asarray(Xs), b200.asarray(Ts))
    if world > 1 and not peer:
        bdist.allreduce_sum_(got)
    got = got.get()
    err = None
    if rank == 0:
        parts = [w.convnet_data(64, n, seed=100 + r)[1:] for r in range(world)]
        with b200.host_arrays():
            want = g(W, onp.concatenate([p_[0] for p_ in parts]), onp.concatenate([p_[1] for p_ in parts]), 1.0)
        err = float(onp.max(onp.abs(got - want)) / onp.max(onp.abs(want)))
    return {"max_rel_err_vs_reference_at_64_images_per_rank": err,
            "workload": f"C5 convnet grad, {N} images split over {world} GPUs ({per} each), 355 KB f64 gradient "
                        f"allreduce per step", "collective": "one-kernel NVLink peer-memory allreduce inside the replayed "
                        "CUDA graph" if peer else ("ncclAllReduce" if world > 1 else "none (1 GPU)"),
            "scaling": "strong", "ms_per_step": ms,
            "grad_evals_per_sec": 1e3 / ms, "images_per_sec": per * world * 1e3 / ms, "launches_per_step": launches}


def run_dp(be, world, rank):
    res = {}
    for name, fn in (("c1_dp", c1_dp), ("c5_dp", c5_dp)):
        try:
            res[name] = fn(be, world, rank)
        except Exception as e:
            res[name] = {"error": f"{type(e).__name__}: {e}"}
        be.synchronize()
        be.empty_cache()
    return res


def run(be, dist=None, which=("kernels", "c1", "c5", "c3", "c4")):
    res = {}
    table = {"kernels": kernels, "c1": c1_mlp, "c3": c3_lstm, "c5": c5_convnet,
             "c4": c4_fluid}
    for name in which:
        try:
            res[name] = table[name](be)
        except Exception as e:  # report, never hide: a failed extra is part of the result
            res[name] = {"error": f"{type(e).__name__}: {e}"}
        be.synchronize()
        be.empty_cache()
    return res
