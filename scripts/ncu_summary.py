"""Condense an .ncu-rep (--set full) into one CSV row per launch with the roofline-relevant metrics."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
labels = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
cols = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(c) for c in cols if c in hdr]
names = {}
if labels:
    for line in open(labels):
        if line.startswith("launches "):
            rng, name = line[len("launches "):].split(":", 1)
            a, b = rng.split("..")
            for k in range(int(a), int(b) + 1):
                names[k] = name.strip()
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "what"] + [hdr[i] for i in idx])
    w.writerow(["", ""] + [rows[1][i] for i in idx])
    for k, r in enumerate(rows[2:]):
        w.writerow([k, names.get(k, "")] + [r[i][:60] for i in idx])
print("wrote", out, len(rows) - 2, "launches")
