"""Per-kernel SASS opcode summary of librife_b200.so: proves which kernels are tcgen05 / TMEM / TMA kernels.
    python tools/sass_summary.py > profiles/r02_sass_summary.md
UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,
SYNCS = mbarrier ops, HFMA2 / FFMA = SIMT math (B200_PROFILING.md)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "eccv2022-rife_b200", "csrc", "librife_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "SHFL", "HFMA2", "FFMA", "LDG", "STG", "LDS", "STS")
kern, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"rife::\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*$", "", kern)
        cnt[kern] = collections.Counter()
        continue
    if kern:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            cnt[kern][op] += 1
            cnt[kern]["_all"] += 1
print("# SASS opcode counts per kernel of librife_b200.so (sm_100a), `python tools/sass_summary.py`\n")
print("| kernel | instr | " + " | ".join(ops) + " |")
print("|---|---|" + "---|" * len(ops))
tot = collections.Counter()
for k, c in cnt.items():
    for o in ops:
        tot[o] += c[o]
    print(f"| `{k[:110]}` | {c['_all']} | " + " | ".join(str(c[o]) for o in ops) + " |")
print(f"| **total** |  | " + " | ".join(str(tot[o]) for o in ops) + " |")
