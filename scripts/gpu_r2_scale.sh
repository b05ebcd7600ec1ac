#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_c2.log 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c2.log").read().strip().splitlines()[-1])
print("C2", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["fp64_instructions_per_point"], d["roofline"]["fp64_issue_frac"], d["parity_max_rel_err_vs_reference"])
PY
{
python scripts/replay_config.py c1 0 200
python scripts/replay_config.py c3 256 3
python scripts/replay_config.py c5 65536 5
} > gpurun_out/cfg.log 2>&1
grep -v Warn gpurun_out/cfg.log
