"""Summarise an `ncu --metrics ... --csv --log-file X.csv` launch list (long format): per-kernel totals and the
slowest launches.  python tools/ncu_launch_summary.py X.csv [top_n]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if r[idx["ID"]] == "ID":
        continue
    d = per.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]]})
    d[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])


def val(d, k, scale=None):
    if k not in d:
        return float("nan")
    v, u = d[k]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return float("nan")
    if scale == "us":
        x *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    if scale == "MB":
        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    return x


def short(n):
    return n.replace("void ", "").replace("rife::", "").replace("<unnamed>::", "").split("(")[0]


agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
lines = []
for k, d in per.items():
    t = val(d, "gpu__time_duration.sum", "us")
    tot += t
    a = agg[short(d["name"]).split("<")[0]]
    a[0] += 1
    a[1] += t
    rd, wr = val(d, "dram__bytes_read.sum", "MB"), val(d, "dram__bytes_write.sum", "MB")
    lines.append((k, short(d["name"])[:44], t, rd, wr, (rd + wr) / t * 1e3 if t else 0,
                  val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  val(d, "sm__warps_active.avg.pct_of_peak_sustained_active"), d.get("launch__grid_size", ("", ""))[0]))
print(f"{len(per)} launches, {tot:.1f} us in kernels (ncu: cold cache, serialised -- compare shares)\n")
print("| kernel | launches | time us | share % |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {t:.1f} | {100 * t / tot:.1f} |")
print(f"\nslowest {top} launches\n\n| id | kernel | time us | dram rd MB | dram wr MB | dram GB/s | tensor pipe active % | warps active % | grid |")
print("|---|---|---|---|---|---|---|---|---|")
for l in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"| {l[0]} | {l[1]} | {l[2]:.1f} | {l[3]:.1f} | {l[4]:.1f} | {l[5]:.0f} | {l[6]:.1f} | {l[7]:.1f} | {l[8]} |")
