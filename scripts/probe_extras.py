"""Development probe: run selected extras at chosen sizes on the GPU and print JSON."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import autograd_b200 as b200
b200.install()
from benchmarks import extras
be = b200.backend.get()
spec = sys.argv[1:]
for s in spec:
    name, *kv = s.split(":")
    kw = {k: int(v) for k, v in (x.split("=") for x in kv)}
    fn = {"kernels": extras.kernels, "c1": extras.c1_mlp, "c3": extras.c3_lstm, "c4": extras.c4_fluid, "c5": extras.c5_convnet}[name]
    try:
        r = fn(be, **kw)
    except Exception as e:
        import traceback; traceback.print_exc()
        r = {"error": repr(e)}
    print(name, kw, json.dumps(r, indent=1), flush=True)
    be.synchronize(); be.empty_cache()
