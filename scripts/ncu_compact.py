"""Reduce an `ncu --page raw --csv` dump to the columns the roofline notes need (one row per profiled launch)."""
import csv
import sys

KEEP = ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
cols = [i for i, n in enumerate(names) if n in KEEP or any(n.startswith(k) for k in ("sm__pipe_tensor", "dram__bytes", "dram__throughput"))]
w = csv.writer(sys.stdout)
w.writerow([names[i] for i in cols])
w.writerow([units[i] for i in cols])
for r in rows[hdr + 2:]:
    if len(r) == len(names):
        w.writerow([r[i][:110] for i in cols])
