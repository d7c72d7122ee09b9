"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and average
duration per kernel.  --last-of N keeps the last 1/N of the launches (the last of N identical iterations)."""
import argparse
import csv
import re
from collections import OrderedDict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--last-of", type=int, default=1)
ap.add_argument("--title", default="")
ap.add_argument("--exclude", default="", help="drop kernels whose name contains this substring (e.g. the spin kernel of bench.py's roofline pass)")
a = ap.parse_args()
rows = [r for r in csv.reader(open(a.csv)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
data = []
for r in rows[1:]:
    try:
        data.append((r[ki], float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)))
    except ValueError:
        pass
if a.exclude:
    data = [(k, v) for k, v in data if a.exclude not in k]
if a.last_of > 1:
    n = len(data) // a.last_of
    data = data[len(data) - n:]
agg = OrderedDict()
for k, v in data:
    k = re.sub(r"\(.*", "", k.replace("void ", "").replace("vlk::<unnamed>::", "").replace("(anonymous namespace)::", ""))[:86]
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
if a.title:
    print(f"# {a.title}\n")
print(f"{len(data)} launches, {tot / 1e3:.2f} ms in total under ncu (cold-cache, serialised: compare SHARES).  "
      f"`at::` kernels: {sum(c for k, (c, _) in agg.items() if (k.startswith('at::') or k.startswith('native::')))} launches, "
      f"{sum(t for k, (_, t) in agg.items() if (k.startswith('at::') or k.startswith('native::'))) / tot * 100:.2f} % of the time.\n")
print("| kernel | launches | total µs | share | avg µs |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {t:.1f} | {t / tot * 100:.1f} % | {t / c:.1f} |")
