#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
} > gpurun_out/gemm_ab.log 2>&1
cat gpurun_out/gemm_ab.log | grep -v Warn
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("bench_c2", "bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], json.dumps(d.get("roofline", {}))[:900], json.dumps(d["e2e"])[:900], json.dumps(d["cpu_baseline"])[:500])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_c2.err
