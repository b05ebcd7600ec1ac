"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the second
half of the launches (= the measured call after the warm-up call)."""
import csv, sys, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    for r in csv.reader(f):
        if len(r) > 14 and r[0].isdigit():
            rows.append((r[4], r[7], r[8], float(r[14])))
half = rows[len(rows) // 2:]
tot = sum(t for *_, t in half)
agg = collections.defaultdict(lambda: [0, 0.0])
for name, blk, grd, t in half:
    key = name.split("(")[0][:60]
    agg[key][0] += 1
    agg[key][1] += t
print(f"{path}: {len(half)} launches, {tot/1e6:.3f} ms total (cold-cache, serialised)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"  {100*t/tot:5.1f}%  {t/1e3:9.1f} us  x{c:<4d} {k}")
