#!/usr/bin/env python
"""bench.py — grad() evaluations per second on the BASELINE.json headline workload.

Workload (config.workload = "C2"): BASELINE.json configs[1] — elementwise_grad nested 4x of tanh
(examples/tanh.py:22-33) over a 2^28-point float64 linspace.  One "step" = one evaluation of
egrad^4(tanh)(x) on one batch (the 2 GiB vector), i.e. forward trace + four nested backward passes through
unmodified autograd, executed by the device backend as one generated fused kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n POINTS] [--extras]

N > 1 is launched by torchrun (one rank per GPU); the workload is elementwise, so ranks take
independent 2^28-point chunks (weak scaling, no data-path collective); timing is barrier + sync
bracketed CUDA events, max over ranks.  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as onp  # noqa: E402

METRIC = "grad_evals_per_sec"
UNIT = "grad evals/s"
ORDER = 4
ALGO_BYTES_PER_POINT = 16  # fused lower bound: read x once + write d4y/dx4 once


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def _dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, local, world, dist


def _bind_to_gpu_numa(gpu_index: int):
    """Run this rank (and so its pinned staging buffers, first-touched by it) on the CPU cores of the GPU's own NUMA
    node, so host<->device copies do not cross the socket interconnect.  Returns what was done, for the JSON line."""
    info = {"numa_node": None, "cpus": None}
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                                   # sysfs uses a 4-digit PCI domain
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = f"{min(cpus)}-{max(cpus)} ({len(cpus)})"
    except Exception as exc:
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


def _copy_ceiling_ms(be, x_host, y_host, dist, reps=3, chunk_bytes=64 << 20):
    """What the box allows: plain pinned cudaMemcpyAsync of the step's input (H2D) and output (D2H) on two streams at
    once, one call per 64 MiB chunk, nothing computed.  Max over ranks; e2e cannot be faster than this."""
    from autograd_b200.lazy import Mat

    n = x_host.size
    din, dout = Mat.empty((n,), x_host.dtype), Mat.empty((n,), y_host.dtype)
    up, down = be.stream_create(), be.stream_create()
    e0, e1, eu, ed = (be.event_create() for _ in range(4))
    step = max(1, chunk_bytes // x_host.itemsize)

    def once():
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            be.h2d_on(din.ptr + lo * x_host.itemsize, x_host[lo:hi], up)
            be.d2h_on(y_host[lo:hi], dout.ptr + lo * y_host.itemsize, down)
        be.event_record_on(eu, up)
        be.event_record_on(ed, down)
        be.stream_wait_event(None, eu)
        be.stream_wait_event(None, ed)

    keep = y_host[:16].copy()
    once()
    be.synchronize()
    _barrier(dist)
    be.event_record(e0)
    for _ in range(reps):
        once()
    be.event_record(e1)
    be.synchronize()
    ms = be.event_elapsed_ms(e0, e1) / reps
    be.stream_destroy(up)
    be.stream_destroy(down)
    del keep
    return _max_over_ranks(ms, dist)


def _barrier(dist):
    if dist is not None:
        dist.barrier()


def _max_over_ranks(value, dist):
    if dist is None:
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------- reference arm / cpu baseline
REF_SAMPLE_POINTS = 1 << 22   # points per timed reference step: 74 live vectors x 32 MiB, far outside any cache


def time_reference_cpu(sample_points: int, full_points: int, steps: int, warmup: int, procs: int | None = None):
    """Reference autograd on NumPy (its own CPU path, unmodified, from baseline/_ref) on a bounded sample of the
    workload: every step evaluates egrad^4(tanh) on `sample_points` points with ALL host cores (one worker process per
    core, each on its own chunk - NumPy's elementwise loops are single-threaded).  Returns whole-workload evals/s
    (the measured points/s divided by the points of one full evaluation), seconds per sample step, cores, description."""
    from benchmarks.cpu_reference import HostPool

    pool = HostPool(procs)
    try:
        for _ in range(max(warmup, 1)):
            pool.step(sample_points)
        times = [pool.step(sample_points) for _ in range(max(steps, 1))]
    finally:
        pool.close()
    dt = sum(times) / len(times)
    value = (sample_points / dt) / full_points
    sample = (f"reference autograd 1.9.1 on NumPy {onp.__version__}: egrad^{ORDER}(tanh) on {sample_points} of "
              f"{full_points} points per step, {pool.procs} worker processes (one per host core, "
              f"{sample_points // pool.procs} points each), {dt:.3f} s per step measured; value = measured points/s "
              f"divided by the points of one full evaluation (elementwise workload: linear in points)")
    return value, dt, pool.procs, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    full = args.n                      # points of ONE evaluation (per GPU in the device arm)
    sample_points = min(full, REF_SAMPLE_POINTS)
    value, dt, cores, sample = time_reference_cpu(sample_points, full, max(args.steps, 1), max(args.warmup, 1))
    # the N-GPU job evaluates N chunks of `full` points per step; the host's points/s does not change with N, so in
    # evaluations of one chunk per second the whole-job figure is the same number
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args.n, args.gpus),
        "step_definition": f"one timed step = {sample_points} points (a bounded sample); one full evaluation = {full} points "
                           f"= {full / sample_points:g} such steps = {dt * full / sample_points:.1f} s at the measured rate",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def _ncu_capture(kernel_name):
    """DRAM traffic and FP64-pipe activity of `kernel_name` from the committed `ncu --set full` captures under profiles/
    (raw-page CSV exports, newest round first).  None when no capture of THIS kernel exists: a stale figure is worse
    than none."""
    import csv
    import glob

    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_c2_fused_ncu_raw*.csv")), reverse=True):
        try:
            with open(path) as f:
                rows = list(csv.reader(f))
            hdr, units = rows[0], rows[1]
            col = {h: i for i, h in enumerate(hdr)}
            for r in rows[2:]:
                if r[col["Kernel Name"]] != kernel_name:
                    continue
                def val(name):
                    return float(r[col[name]].replace(",", "")) * unit_scale.get(units[col[name]], 1.0)
                return {"traffic": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                        "fp64_pipe_active_pct": float(r[col["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
                        "ncu_kernel_ms": float(r[col["gpu__time_duration.sum"]]) *
                        {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0),
                        "source": os.path.relpath(path, ROOT)}
        except Exception:
            continue
    return None


def _config(n, gpus):
    return {"workload": f"C2: elementwise_grad nested {ORDER}x of tanh (examples/tanh.py) over a {n}-point float64 "
                        f"linspace per GPU", "points_per_gpu": n, "order": ORDER,
            "parallelism": f"independent chunks x{gpus}" if gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (2 GiB in, 2 GiB out per step)" if n * 8 > (126 << 20) else
                         "input smaller than L2 (reduced --n run)"}


# ----------------------------------------------------------------------------- device arm
def run_b200(args):
    rank, local, world, dist = _dist_setup(args.gpus)
    os.environ.setdefault("B200AG_DEVICE", str(local))
    full_affinity = os.sched_getaffinity(0)
    numa = _bind_to_gpu_numa(local)       # before any pinned allocation
    import autograd_b200 as b200

    b200.install()
    from examples import workloads as w

    be = b200.backend.get()
    if be.name != "cuda":
        raise SystemExit("bench.py must run on the CUDA backend")
    n = args.n
    # this rank's chunk of the N*n-point linspace (independent elements: no collective on the data path)
    lo = -7.0 + 14.0 * rank / world
    hi = -7.0 + 14.0 * (rank + 1) / world
    x_host = b200.array.pinned_empty((n,), onp.float64)
    x_host[:] = onp.linspace(lo, hi, n, endpoint=(world == 1))
    y_host = b200.array.pinned_empty((n,), onp.float64)
    f = w.tanh_nth_derivative(ORDER)

    # ---- device-resident measurement: inputs already in HBM when the timed region starts
    x = b200.asarray(x_host)
    for _ in range(max(args.warmup, 3)):
        f(x).materialize()
    be.synchronize()
    ev0, ev1 = be.event_create(), be.event_create()
    l0 = be.launch_count()
    with ClockSampler(local) as clocks:
        time.sleep(0.06)
        _barrier(dist)
        be.synchronize()
        be.event_record(ev0)
        for _ in range(args.steps):
            y = f(x).materialize()
            del y
        be.event_record(ev1)
        be.synchronize()
        _barrier(dist)
        launches = be.launch_count() - l0
        eager_ms_local = be.event_elapsed_ms(ev0, ev1) / args.steps
        # ---- the same evaluation replayed as a CUDA graph (autograd_b200.capture): the eager call re-traces the four
        # nested backward passes through autograd's Python tracer on every step (~5 ms of host work, now longer than the
        # kernel), the replay launches the recorded tape — here ONE fused kernel — without it
        replay_ms, replay_nodes = None, None
        try:
            fast = b200.capture(f)
            for _ in range(max(args.warmup, 3)):
                fast(x)
            be.synchronize()         # ranks left the barrier above together; no collective inside this try block
            be.event_record(ev0)
            for _ in range(args.steps):
                fast(x)
            be.event_record(ev1)
            be.synchronize()
            replay_ms = be.event_elapsed_ms(ev0, ev1) / args.steps
            replay_nodes = int(list(fast.num_nodes().values())[0])
            del fast
            be.empty_cache()
        except Exception as exc:   # keep the eager figure, say why
            replay_err = f"{type(exc).__name__}: {exc}"
            print("bench: graph replay of C2 failed:", replay_err, file=sys.stderr)
        if clocks.samples is not None and len(clocks.samples) < 3:
            # a short timed region can fall between two nvidia-smi polls: keep the same kernel running (untimed)
            # until a few samples under load exist
            t_end = time.perf_counter() + 1.0
            while len(clocks.samples) < 3 and time.perf_counter() < t_end:
                f(x).materialize()
                be.synchronize()
    eager_ms = _max_over_ranks(eager_ms_local, dist)
    # every rank takes part in both reductions; one failed capture anywhere disables the replay figure everywhere
    failed = _max_over_ranks(1.0 if replay_ms is None else 0.0, dist)
    replay_ms = _max_over_ranks(replay_ms or 0.0, dist) if not failed else None

    use_replay = replay_ms is not None and replay_ms < eager_ms
    ms_per_step = replay_ms if use_replay else eager_ms
    if use_replay:
        launches = replay_nodes * args.steps
    value = world / (ms_per_step * 1e-3)

    # ---- dominant kernel: per-launch duration on the launching stream (CUDA events around each launch)
    be.profile = []
    for _ in range(args.steps):
        f(x).materialize()
    be.synchronize()
    durs = [(prog, npts, be.event_elapsed_ms(e0, e1)) for prog, npts, e0, e1 in be.profile]
    be.profile = None
    kern_ms = sum(d for _, _, d in durs) / max(len(durs), 1)
    peak, peak_src = _peaks()
    achieved = ALGO_BYTES_PER_POINT * n / (kern_ms * 1e-3) / 1e9
    prog = durs[0][0] if durs else None
    # DRAM traffic per launch and FP64-pipe activity: read from the committed `ncu --set full` capture of THIS kernel
    # (matched by its generated name, which is a hash of the program) - never typed in
    kname, fp64_per_point = None, None
    if prog is not None:
        from autograd_b200 import codegen

        src_, kname, _ = codegen.generate(prog)
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            from count_body_instrs import per_point_counts

            cnt, _sass = per_point_counts(src_)
            fp64_per_point = int(sum(cnt[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")))
        except Exception as exc:      # cuobjdump / NVRTC unavailable: leave the count out
            print("bench: per-point instruction count unavailable:", exc, file=sys.stderr)
    cap = _ncu_capture(kname) if kname else None
    sm_clock_hz = 1.965e9
    # FP64 issue roofline: FP64-pipe instructions per point x points / (SMs x 64 lanes per clock x clock)
    fp64_floor_ms = (fp64_per_point * n / (be.sm_count * 64 * sm_clock_hz) * 1e3) if fp64_per_point else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (cap["traffic"] * n / float(1 << 28)) if cap else None,
                "traffic_source": (f"ncu --set full, {cap['source']} (kernel {kname}, 2^28 points, scaled to --n)" if cap else
                                   f"no committed capture of kernel {kname} under profiles/"),
                "fp64_pipe_active_pct_ncu": cap["fp64_pipe_active_pct"] if cap else None,
                "fp64_instructions_per_point": fp64_per_point,
                "fp64_issue_floor_ms": fp64_floor_ms,
                "fp64_issue_frac": (fp64_floor_ms / kern_ms) if fp64_floor_ms else None,
                "peak_source": peak_src, "kernel": f"fused elementwise {kname} (generated, NVRTC sm_100a)",
                "kernel_ms": kern_ms, "kernel_ops_fused": len(prog.instrs) if prog else None,
                "kernel_share_of_step": kern_ms / ms_per_step,
                "note": "algorithmic bytes = 16 B/point (read x, write y).  The body is 348 recorded ufunc calls; everything "
                        "downstream of exp() is inexact against NumPy anyway, so multiply-add pairs are contracted and "
                        "quotients share one refined reciprocal (bit-exact rounding is kept only on values that can still "
                        "be bit-identical to NumPy's).  The kernel is bound by FP64 issue, not by HBM: fp64_issue_frac is "
                        "the fraction of the FP64-pipe roofline (static SASS count of the body, 64 lanes/clk/SM at 1.965 GHz)"}

    # ---- end to end through the public API with HOST buffers: H2D + grad + D2H inside the timed region
    for _ in range(2):
        f(b200.asarray(x_host)).get(out=y_host)
    be.synchronize()
    e0, e1 = be.event_create(), be.event_create()
    e2e_steps = max(1, min(args.steps, 5))
    _barrier(dist)
    be.synchronize()
    be.event_record(e0)
    for _ in range(e2e_steps):
        f(b200.asarray(x_host)).get(out=y_host)
    be.event_record(e1)
    be.synchronize()
    _barrier(dist)
    e2e_ms = _max_over_ranks(be.event_elapsed_ms(e0, e1) / e2e_steps, dist)
    # the same call through the chunked public API: upload of chunk i+1 / compute of i / download of i-1 overlap
    pipe_ms = None
    try:
        for _ in range(2):
            b200.map_elementwise(f, x_host, out=y_host)
        _barrier(dist)
        be.synchronize()
        be.event_record(e0)
        for _ in range(e2e_steps):
            b200.map_elementwise(f, x_host, out=y_host)
        be.event_record(e1)
        be.synchronize()
        _barrier(dist)
        pipe_ms = _max_over_ranks(be.event_elapsed_ms(e0, e1) / e2e_steps, dist)
    except Exception as exc:  # report, keep the monolithic number
        pipe_ms = None
        pipe_err = f"{type(exc).__name__}: {exc}"
    best_ms = min(e2e_ms, pipe_ms) if pipe_ms else e2e_ms
    # the box's own ceiling for this much host<->device traffic on `world` GPUs at once (nothing computed)
    try:
        ceiling_ms = _copy_ceiling_ms(be, x_host, y_host, dist)
    except Exception as exc:
        ceiling_ms = None
        print("bench: copy ceiling probe failed:", exc, file=sys.stderr)
    e2e = {"value": world / (best_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.nbytes),
           "d2h_bytes_per_step": int(y_host.nbytes), "ms_per_step": best_ms,
           "api": "autograd_b200.map_elementwise(f, x_host, out=y_host)" if pipe_ms and pipe_ms <= e2e_ms else
                  "f(autograd_b200.asarray(x_host)).get(out=y_host)",
           "monolithic_ms_per_step": e2e_ms, "chunk_pipelined_ms_per_step": pipe_ms,
           "copy_ceiling_ms": ceiling_ms,
           "copy_ceiling_note": "plain pinned cudaMemcpyAsync H2D + D2H of the same bytes on two streams, all ranks at once, "
                                "max over ranks, no compute",
           "frac_of_copy_ceiling": (ceiling_ms / best_ms) if ceiling_ms else None,
           "host_binding": numa}

    # parity spot check on the bench inputs themselves: 2^16 points strided across the whole resident vector,
    # error relative to the max-norm of the reference result (the derivative crosses zero)
    stride = max(1, n >> 16)
    got = f(x[::stride]).get()
    with b200.host_arrays():
        want = f(x_host[::stride].copy())
    with b200.host_arrays():
        scale = float(onp.max(onp.abs(f(onp.linspace(-7.0, 7.0, 4097)))))  # max-norm of the function, whole range
    err = float(onp.max(onp.abs(got - want)) / scale)
    mem = be.mem_stats()

    # ---- batch-sharded configs (C1 weak, C5 strong) with the NCCL gradient allreduce: every rank takes part
    dp_results = None
    if not args.no_extras:
        del x, x_host, y_host, got, want
        be.synchronize()
        be.empty_cache()
        from autograd_b200 import dist as bdist
        from benchmarks import extras

        if world > 1:
            bdist.init("nccl")
        dp_results = extras.run_dp(be, world, rank)

    if rank != 0:
        if dist is not None:
            dist.barrier()
        return 0

    # ---- reference CPU path timed beside it on this box's host cores (bounded sample)
    sample_points = min(n, REF_SAMPLE_POINTS)
    os.sched_setaffinity(0, full_affinity)      # the CPU baseline may use every host core again
    cpu_value, cpu_dt, cores, sample = time_reference_cpu(sample_points, n, 3, 1)
    cpu = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
           "seconds_per_sample_step": cpu_dt, "host_cpu_count": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": _config(n, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks.summary(),
        "execution": {"timed": "CUDA-graph replay (autograd_b200.capture)" if use_replay else "eager (autograd re-traces every step)",
                      "eager_ms_per_step": eager_ms, "replay_ms_per_step": replay_ms, "replay_graph_nodes": replay_nodes},
        "parity_max_rel_err_vs_reference": err, "mem": mem,
    }
    if dp_results is not None:
        line["other_configs"] = dp_results
    if not args.no_extras and world == 1:
        from benchmarks import extras

        line["other_configs"] = {**extras.run(be), **(dp_results or {})}
    _emit(line)
    if dist is not None:
        dist.barrier()
    return 0


_REAL_STDOUT = None


def _emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout must carry exactly ONE JSON line; native libraries (NCCL prints its version banner) write to fd 1
    # directly, so everything else is sent to stderr and the JSON goes to a private duplicate of the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1 << 28, help="points per GPU")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary measurements (configs C1/C3/C4/C5, kernel rates) reported under other_configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
