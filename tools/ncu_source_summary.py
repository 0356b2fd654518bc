"""Summarise an ncu --page source --csv dump: top SASS instructions by stall samples, plus per-reason totals."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = {h: 0 for h in stall_cols}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: ns = int(r[idx["# Samples"]])
    except Exception: continue
    ie = r[idx["Instructions Executed"]]
    st = {h: int(r[idx[h]] or 0) for h in stall_cols}
    for h in stall_cols: tot[h] += st[h]
    data.append((ns, r[idx["Source"]][:100], ie, st))
total = sum(d[0] for d in data)
print("total samples", total)
print("stall totals:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.01 * total})
print("total warp-instructions executed:", sum(int(d[2] or 0) for d in data))
for ns, src, ie, st in sorted(data, key=lambda d: -d[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%6d %5.1f%% inst=%9s  %-90s %s" % (ns, 100.0 * ns / total, ie, src, top))
