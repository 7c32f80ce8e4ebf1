"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump of ONE inference pass (tools/ncu_target.py) into the
per-launch table kept under profiles/:  python tools/ncu_summary.py raw.csv[.gz] hdv3|eccv > profiles/rNN_....md"""
import csv
import gzip
import sys

path, mode = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "hdv3")
op = gzip.open if path.endswith(".gz") else open
lines = op(path, "rt").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))   # skip ncu's ==PROF== preamble of --log-file dumps
rows = list(csv.reader(lines[start:]))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[idx[k]].replace(",", ""))
    except (ValueError, KeyError):
        return float("nan")


names = []
if mode == "hdv3" and len(data) == 65:   # block 4's input gather fused into its conv0.0 (conv_gather.cu)
    names = ["prep"] + ["head.cnn%d" % i for i in range(4)]
    for b in range(5):
        names += ([f"b{b}.input+conv0.0"] if b == 4 else [f"b{b}.input", f"b{b}.conv0.0"]) + [f"b{b}.conv0.1"]
        names += [f"b{b}.res{j}" for j in range(8)] + [f"b{b}.lastconv"]
    names += ["merge"]
elif mode == "hdv3" and len(data) == 66:
    names = ["prep"] + ["head.cnn%d" % i for i in range(4)]
    for b in range(5):
        names += [f"b{b}.input", f"b{b}.conv0.0", f"b{b}.conv0.1"] + [f"b{b}.res{j}" for j in range(8)] + [f"b{b}.lastconv"]
    names += ["merge"]
tu, bu = units[idx["gpu__time_duration.sum"]], units[idx["dram__bytes_read.sum"]]
scale_b = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[bu]
scale_t = {"us": 1.0, "ms": 1e3, "ns": 1e-3}[tu]
print("| # | stage | kernel | time us | dram rd MB | dram wr MB | dram GB/s | dram % of ncu peak | tensor pipe active % | warps active % | regs | grid | L2 hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
agg = {}
for i, r in enumerate(data):
    t = num(r, "gpu__time_duration.sum") * scale_t
    rd, wr = num(r, "dram__bytes_read.sum") * scale_b, num(r, "dram__bytes_write.sum") * scale_b
    k = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:40]
    st = names[i] if i < len(names) else ""
    tot += t
    a = agg.setdefault(k.split("<")[0], [0, 0.0])
    a[0] += 1
    a[1] += t
    print(f"| {i} | {st} | {k} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t / 1e3:.0f} | "
          f"{num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {r[idx['launch__registers_per_thread']]} | "
          f"{r[idx['launch__grid_size']]} | {num(r, 'lts__t_sector_hit_rate.pct'):.1f} |")
print(f"\ntotal kernel time {tot:.1f} us over {len(data)} launches (ncu: cold cache, serialised -- compare shares)\n")
print("| kernel | launches | time us | share % |")
print("|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f} |")
