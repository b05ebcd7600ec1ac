#!/bin/bash
# round 2, first GPU pass: full GPU suite, smoke, C2 exact-vs-relaxed A/B, bench with extras, launch list + full ncu of the C2 kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log; echo "tests took $(( $(date +%s) - t0 )) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
B200AG_EXACT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_c2_exact.log 2> gpurun_out/bench_c2_exact.err; echo "exact rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_c2_relaxed.log 2> gpurun_out/bench_c2_relaxed.err; echo "relaxed rc=$?"
t1=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; echo "bench took $(( $(date +%s) - t1 )) s"
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c2_launches.csv $CMD > gpurun_out/c2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_ -s 4 -c 1 -o gpurun_out/r02_prof_c2 $CMD > gpurun_out/c2_ncu_full.log 2>&1
tail -1 gpurun_out/c2_ncu_full.log
python - <<'PY'
import json
for f in ("bench_c2_exact", "bench_c2_relaxed", "bench_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["parity_max_rel_err_vs_reference"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
