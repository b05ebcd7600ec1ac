#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_config.py c4 2 > gpurun_out/plain_c4.log 2>&1 || exit 1
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none -s 124 -c 124 --csv --log-file gpurun_out/r02_c4_sol.csv python scripts/profile_config.py c4 2 > gpurun_out/ncu_c4_sol.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_c4_sol.csv")))
hdr = None
data = collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        data.setdefault((d["ID"], d["Kernel Name"][:40]), {})[d["Metric Name"]] = d["Metric Value"]
agg = collections.OrderedDict()
for (kid, name), m in data.items():
    def f(k):
        try: return float(m.get(k, "nan").replace(",", ""))
        except: return float("nan")
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += f("Duration"); a[2] += f("DRAM Throughput"); a[3] += f("L2 Cache Throughput"); a[4] += f("Compute (SM) Throughput"); a[5] += f("Memory Throughput")
print("kernel                                    calls  total_us  avg_us  DRAM%  L2%  SM%  Mem%")
for name, (c, dur, dram, l2, sm, mem) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:42s} {c:4d} {dur:9.1f} {dur/c:7.1f} {dram/c:6.1f} {l2/c:5.1f} {sm/c:5.1f} {mem/c:5.1f}")
PY
