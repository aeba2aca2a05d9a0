#!/usr/bin/env python
"""Static SASS opcode histogram of libenf_b200.so per kernel family (cuobjdump -sass): the evidence that the
tensor-core / TMA kernels are Blackwell-native (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG =
TMA tensor load/store, UBLKCP = cp.async.bulk, FFMA2/FMUL2/FADD2 = packed FP32 pairs, MUFU.* = special-function unit).

  python tools/sass_hist.py [path/to/libenf_b200.so] > profiles/r2_sass_opcodes.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "euclidiannormalizingflows.jl_b200", "libenf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA",
         "DFMA", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "MUFU.RSQ", "MUFU.SQRT", "MUFU.RCP64H", "MUFU.RSQ64H", "SHFL", "LDS", "STS",
         "LDG", "STG", "RED", "ATOM"]
fam = collections.defaultdict(collections.Counter)
nk = collections.Counter()
arch = set()
cur = None
for ln in out.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        key = re.sub(r"<.*", "", d.split("(")[0]).split("::")[-1].replace("void ", "")
        cur = key
        nk[key] += 1
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        fam[cur]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                fam[cur][w] += 1
print(f"{os.path.basename(so)}: {os.path.getsize(so) / 1e6:.1f} MB, arch {sorted(arch)}, {sum(nk.values())} kernels in {len(nk)} families")
tot = collections.Counter()
for k in sorted(fam, key=lambda k: -fam[k]["total"]):
    c = fam[k]
    print(f"\n{k}  ({nk[k]} instantiations, {c['total']} SASS instructions)")
    print("   " + "  ".join(f"{w}:{c[w]}" for w in WATCH if c[w]))
    tot.update(c)
print("\nwhole library: " + "  ".join(f"{w}:{tot[w]}" for w in WATCH if tot[w]))
