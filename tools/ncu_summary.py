"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
one training step (tools/profile_one_step.py) into
  * a compact per-launch CSV (id,kernel,time_us,dram_read_bytes,dram_write_bytes),
  * a per-kernel summary table,
  * the entry of profiles/r2_traffic.json that bench.py reports as `roofline.traffic`.
usage: ncu_summary.py raw.csv out_prefix [config batch]"""
import csv, json, os, re, subprocess, sys
raw, prefix = sys.argv[1], sys.argv[2]
config = sys.argv[3] if len(sys.argv) > 3 else "ukbb192"
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 128
rows = {}
with open(raw) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    e = rows.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        e["time_us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif r["Metric Name"] == "dram__bytes_read.sum":
        e["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif r["Metric Name"] == "dram__bytes_write.sum":
        e["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(k):
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"<unnamed>::", "", k)
    k = re.sub(r"\(.*$", "", k)
    return k.replace(",", ";")[:48]


with open(prefix + "_launches.csv", "w") as f:
    f.write(f"# ncu launch list of one eager training step: {config} batch {batch}\nid,kernel,time_us,dram_read_bytes,dram_write_bytes\n")
    for i in sorted(rows):
        e = rows[i]
        f.write("%d,%s,%.2f,%d,%d\n" % (i, short(e["kernel"]), e.get("time_us", 0), e.get("rd", 0), e.get("wr", 0)))
agg = {}
for e in rows.values():
    a = agg.setdefault(short(e["kernel"]), [0, 0.0, 0.0])
    a[0] += 1; a[1] += e.get("time_us", 0); a[2] += e.get("rd", 0) + e.get("wr", 0)
tot = sum(a[1] for a in agg.values())
with open(prefix + "_summary.txt", "w") as f:
    f.write("launches %d total %.0f us\n%-50s %5s %10s %8s %6s %s\n" % (len(rows), tot, "kernel", "n", "us", "avg us", "share", "dram MB/launch"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-50s %5d %10.0f %8.1f %6.3f %14.2f\n" % (k, a[0], a[1], a[1] / a[0], a[1] / tot, a[2] / a[0] / 1e6))
print(open(prefix + "_summary.txt").read())
conv = [a for k, a in agg.items() if "conv_tc_kernel" in k]
if conv:
    n = sum(a[0] for a in conv); by = sum(a[2] for a in conv); us = sum(a[1] for a in conv)
    tj = os.path.join(os.path.dirname(os.path.abspath(prefix)), "r2_traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    try:
        commit = subprocess.check_output(["git", "rev-parse", "--short", "HEAD"], text=True).strip()
    except Exception:
        commit = None
    d[f"{config}:{batch}"] = {"conv_dram_bytes_per_launch": by / n, "conv_launches": n, "conv_time_us_serialised": us,
                             "conv_share_of_serialised_step": us / tot,
                             "source": os.path.basename(prefix) + "_launches.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                             "summarised_at_commit": commit}
    json.dump(d, open(tj, "w"), indent=1)
    print("traffic entry:", d[f"{config}:{batch}"])
