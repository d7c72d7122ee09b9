"""profiles/rNN_ncu_gemm_shapes.json from the compact ncu capture of `kernel_zoo.py --once --only "gemm CLIP"` (the order of the
launches is the order of the zoo's CLIP cases): DRAM bytes, duration and tensor-pipe share per launch shape — bench.py reads
the DRAM bytes of its dominant kernel (roofline.traffic) from this file."""
import csv
import json
import sys

CASES = [([16448, 3072, 1024], "lnfold", "folded LayerNorm + bias"),
         ([16448, 4096, 1024], "lnfold", "folded LayerNorm + bias + quick_gelu"),
         ([16448, 3072, 1024], "gemm", "bias"),
         ([16448, 1024, 1024], "gemm", "bias + residual"),
         ([16448, 4096, 1024], "gemm", "bias + quick_gelu"),
         ([16448, 1024, 4096], "gemm", "bias + residual")]
rows = list(csv.reader(open(sys.argv[1])))
names, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(names)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
gemm_rows = [r for r in rows[2:] if "gemm_bf16" in r[idx["Kernel Name"]]]
# --once launches every case twice (warm-up + the profiled one): keep the second of each pair
if len(gemm_rows) == 2 * len(CASES):
    gemm_rows = gemm_rows[1::2]
out = []
for (shape, kind, epi), r in zip(CASES, gemm_rows):
    f = lambda n: float(r[idx[n]].replace(",", ""))
    out.append({"shape": shape, "kind": kind, "epilogue": epi,
                "time_us": f("gpu__time_duration.sum") * tscale.get(units[idx["gpu__time_duration.sum"]], 1.0),
                "dram_bytes_read": f("dram__bytes_read.sum") * scale[units[idx["dram__bytes_read.sum"]]],
                "dram_bytes_write": f("dram__bytes_write.sum") * scale[units[idx["dram__bytes_write.sum"]]],
                "tensor_pipe_pct_elapsed": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")})
json.dump({"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "shapes": out}, sys.stdout, indent=1)
