"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...` launch list:
    python scripts/launch_summary.py gpurun_out/launches.csv > profiles/rN_launches_summary.txt
Times are cold-cache, serialised per-launch durations: the SHARE of each kernel is what compares with the live bench."""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    t = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    n, tot = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, tot + t)
total = sum(t for _, t in agg.values())
print(f"# total {total / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {n:8d} {t:12.1f} {100 * t / total:6.1f}%")
