#!/usr/bin/env python3
"""Dynamic instruction mix and stall samples per opcode from `ncu -i X.ncu-rep --page source --csv` (SASS view).

    ncu -i gpurun_out/X.ncu-rep --page source --csv > /tmp/src.csv ; python tools/ncu_mix.py /tmp/src.csv
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
ex, smp, reuse = defaultdict(float), defaultdict(float), defaultdict(float)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = defaultdict(lambda: defaultdict(float))
tot = 0.0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    try:
        n = float(r[col["Instructions Executed"]])
        s = float(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    ex[op] += n
    smp[op] += s
    tot += n
    if ".reuse" in src:
        reuse[op] += n
    for h in stall_cols:
        try:
            stalls[op][h] += float(r[col[h]] or 0)
        except ValueError:
            pass
ts = sum(smp.values())
print(f"total warp instructions {tot:.4g}, samples {ts:.0f}")
print(f"{'op':10s} {'executed%':>9s} {'samples%':>9s}  top stalls (% of all samples)")
for op, n in sorted(ex.items(), key=lambda kv: -kv[1])[:22]:
    st = sorted(stalls[op].items(), key=lambda kv: -kv[1])[:3]
    print(f"{op:10s} {100 * n / tot:9.2f} {100 * smp[op] / ts if ts else 0:9.2f}  " +
          ", ".join(f"{k[6:]} {100 * v / ts:.1f}" for k, v in st if v))
