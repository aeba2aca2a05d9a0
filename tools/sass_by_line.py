#!/usr/bin/env python
"""Source-level hot spots of one kernel: join the per-SASS-instruction execution counts of an ncu report with the
line table of the cubin (nvdisasm --print-line-info), aggregate by (file, line) and by inlined function.

  python tools/sass_by_line.py REPORT.ncu-rep OBJECT.o MANGLED_KERNEL_NAME [units]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, fun = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
elf = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
idx = None
for ln in elf.split("\n"):
    if ln.rstrip().endswith(" .text." + fun) and "PROGBITS" in ln:
        idx = ln.split()[-2]                      # section info field = symbol index of the function (hex)
assert idx is not None, "kernel not found in " + obj
dis = subprocess.run(["nvdisasm", "--print-line-info", "-fun", "0x" + idx, cubin], capture_output=True, text=True).stdout
i0 = dis.index("\t.section\t.text." + fun)
i1 = dis.find("\t.section\t", i0 + 10)
dis = dis[i0:i1 if i1 > 0 else len(dis)]
# walk the listing: remember the current line annotation, attach it to every instruction (/*addr*/)
cur = ("?", 0)
line_of = []
for ln in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ie, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
assert len(data) == len(line_of), (len(data), len(line_of))
by_line, samp_line, ops_line = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
tot = 0
for r, (addr, loc, text) in zip(data, line_of):
    n = int(r[ie]); tot += n
    by_line[loc] += n
    samp_line[loc] += int(r[isamp])
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ops_line[loc][op] += n
ts = sum(samp_line.values()) or 1
print(f"total warp instructions {tot}" + (f" = {tot / units:.2f}/unit" if units else ""))
for loc, n in by_line.most_common(70):
    per = f"{n / units:7.2f}/unit" if units else f"{100 * n / tot:5.1f}%"
    mix = " ".join(f"{k}:{v / units:.2f}" if units else f"{k}:{v}" for k, v in ops_line[loc].most_common(6))
    print(f"{loc[0]:18s}:{loc[1]:5d} {per}  stall {100 * samp_line[loc] / ts:5.1f}%   {mix}")
