"""Markdown table of the compact ncu CSVs (scripts/ncu_compact.py output): one row per profiled launch."""
import csv
import glob
import os
import sys

files = sorted(glob.glob(os.path.join(sys.argv[1], "ncu_*.csv")))
print("| capture | kernel | grid x block | regs | time us | DRAM read MB | DRAM write MB | DRAM thr % | achieved GB/s | tensor pipe % (elapsed) | L2 thr % | SM thr % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for f in files:
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    names, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(names)}
    unit = {n: units[i] for i, n in enumerate(names)}

    def val(r, n, default=0.0):
        try:
            return float(r[idx[n]].replace(",", ""))
        except Exception:
            return default

    def to_mb(r, n):
        v, u = val(r, n), unit.get(n, "")
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)

    def to_us(r, n):
        v, u = val(r, n), unit.get(n, "")
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)

    for r in rows[2:]:
        k = r[idx["Kernel Name"]].replace("void ", "").replace("unnamed>::", "").replace("vlk::<", "")
        k = k.split("(")[0][:52]
        t = to_us(r, "gpu__time_duration.sum")
        rd, wr = to_mb(r, "dram__bytes_read.sum"), to_mb(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / t * 1e3 if t > 0 else 0
        print(f"| {os.path.basename(f)[4:-4]} | `{k}` | {r[idx['Grid Size']]} x {r[idx['Block Size']]} | "
              f"{int(val(r, 'launch__registers_per_thread'))} | {t:.1f} | {rd:.1f} | {wr:.1f} | "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {gbs:.0f} | "
              f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
