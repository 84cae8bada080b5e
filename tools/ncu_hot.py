"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples and the
stall-reason totals.  usage: python tools/ncu_hot.py src.csv [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
S = col["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
print("kernel:", rows[0][1] if rows[0] else "?", " total samples", tot, " instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stalls}
print("stall totals:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in
                                 sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
ie = col["Instructions Executed"]
print("warp instructions executed:", sum(int(r[ie] or 0) for r in data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:top]
for i in sorted(order):
    r = data[i]
    main = max(stalls, key=lambda h: int(r[col[h]] or 0))
    print("%5d %6.2f%% exec=%-9s %-12s %s" % (i, 100.0 * int(r[S] or 0) / max(tot, 1), r[ie], main[6:],
                                              r[col["Source"]][:110]))
