#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "row_mode or roll or views or fluid or c4 or max_pool or convnet or c5 or short_axis or broadcasting or lstm" > gpurun_out/rowchunk_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/rowchunk_tests.log
timeout 600 python bench.py --steps 10 > gpurun_out/rowchunk_bench.log 2> gpurun_out/rowchunk_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/rowchunk_bench.log') if l.startswith('{')][-1])
oc=d['other_configs']
for k,v in oc['kernels'].items():
    if 'GB/s' in v: print(f"{k:70s} {v['ms']:.4f} {v['frac_of_hbm_peak']:.2f}")
for k in ('c1','c3','c4','c5'):
    print(k, oc[k].get('ms'), (oc[k].get('graph_replay') or {}).get('ms'))
PY
