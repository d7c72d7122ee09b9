"""Summarise an `ncu --page source --csv` dump (SASS view): stall-reason totals and the hottest instructions."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
idx = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = Counter()
inst = []
opc = Counter()
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    n = int(r[idx["# Samples"]] or 0)
    for c in stall_cols:
        tot[c] += int(r[idx[c]] or 0)
    op = r[idx["Source"]].strip().split()[0] if r[idx["Source"]].strip() else "?"
    if op.startswith("@"):
        op = r[idx["Source"]].strip().split()[1]
    opc[op.split(".")[0]] += n
    inst.append((n, r[idx["Source"]].strip()[:90], {c: int(r[idx[c]] or 0) for c in stall_cols if int(r[idx[c]] or 0) > 0}))
total = sum(x[0] for x in inst)
print("total samples", total)
print("stall reasons:", ", ".join(f"{k[6:]} {100*v/max(total,1):.1f}%" for k, v in tot.most_common(10)))
print("by opcode:", ", ".join(f"{k} {100*v/max(total,1):.1f}%" for k, v in opc.most_common(14)))
print("hottest instructions:")
for n, src, st in sorted(inst, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"  {100*n/max(total,1):5.1f}%  {src:90s} {', '.join(f'{k[6:]}:{v}' for k, v in top)}")
