"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_summary.py profiles/launches_r02_bench.csv > profiles/launches_r02_bench_summary.csv
Columns: kernel, launches, total_us, avg_us, share (of the summed kernel time of the list)."""
import csv
import sys
from collections import OrderedDict

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
acc = OrderedDict()
for d in csv.DictReader(rows):
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    ns = float(d["Metric Value"].replace(",", ""))
    if d.get("Metric Unit") in ("us", "usecond"):
        ns *= 1e3
    n, t = acc.get(d["Kernel Name"], (0, 0.0))
    acc[d["Kernel Name"]] = (n + 1, t + ns)
total = sum(t for _, t in acc.values()) or 1.0
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "avg_us", "share"])
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, n, f"{t / 1e3:.1f}", f"{t / 1e3 / n:.2f}", f"{t / total:.4f}"])
