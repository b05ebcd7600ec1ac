"""Time the CUDA-graph replay of one BASELINE config (optionally at reduced length) — for knob A/B runs.

    python scripts/replay_config.py c4 [steps=20] [reps=5]      -> "c4 steps=20 replay_ms=... nodes=..."
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import autograd_b200 as b200
b200.install()
from autograd import grad, value_and_grad
from benchmarks.extras import tree_to_dev
from examples import workloads as w

be = b200.backend.get()
cfg = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
if cfg == "c1":
    params, X, T = w.mlp_data(); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(T), 1.0); f = grad(w.mlp_objective)
elif cfg == "c3":
    params, X, Y = w.lstm_data(seq_len=size or 16); a = (tree_to_dev(params), b200.asarray(X), b200.asarray(X)); f = grad(w.lstm_objective)
elif cfg == "c4":
    p, s0, tg = w.fluid_data(2048, 2048); a = (b200.asarray(p), b200.asarray(s0), b200.asarray(tg), 2048, 2048, size or 20)
    f = value_and_grad(w.fluid_objective)
elif cfg == "c5":
    n, _p, loss = w.convnet_make(); W, X, T = w.convnet_data(size or 65536, n); a = (b200.asarray(W), b200.asarray(X), b200.asarray(T)); f = grad(loss)
elif cfg == "c2":
    import numpy as onp
    from autograd import elementwise_grad as egrad
    import autograd.numpy as np
    f = egrad(egrad(egrad(egrad(np.tanh))))
    a = (b200.asarray(onp.linspace(-7.0, 7.0, size or (1 << 28))),)
else:
    raise SystemExit("config must be c1 | c2 | c3 | c4 | c5")
fast = b200.capture(f)
fast(*a); fast(*a); be.synchronize()
e0, e1 = be.event_create(), be.event_create()
be.event_record(e0)
for _ in range(reps):
    fast(*a)
be.event_record(e1); be.synchronize()
knobs = {k: v for k, v in os.environ.items() if k.startswith("B200AG_")}
print(f"{cfg} size={size} replay_ms={be.event_elapsed_ms(e0, e1) / reps:.3f} nodes={list(fast.num_nodes().values())[0]} knobs={knobs}")
