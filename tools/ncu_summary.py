#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the text summary
kept under profiles/: headline metrics, stall reasons, dynamic instruction mix
and the SASS lines with the most stall samples.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--units N_WARP_ELEMENTS] > profiles/xyz.txt
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    units = None
    if "--units" in sys.argv:
        units = float(sys.argv[sys.argv.index("--units") + 1])
    raw = page(rep, "raw")
    h, v = raw[0], raw[2]
    get = lambda k: v[h.index(k)] if k in h else "n/a"
    print("kernel:", get("Kernel Name")[:150])
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for k in keys:
        if k in h:
            print(f"  {k:70s} {get(k)} {raw[1][h.index(k)]}")
    print("stall reasons (warps per issue-active cycle):")
    for k in h:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            val = float(get(k))
            if val >= 0.05:
                print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {val:6.2f}")
    src = page(rep, "source")
    if len(src) < 3:
        return
    hdr, data = src[1], src[2:]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in data)
    ts = sum(int(r[isamp]) for r in data) or 1
    by, samp = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ia].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        by[op] += int(r[ie])
        samp[op] += int(r[isamp])
    print(f"dynamic instruction mix: {tot} warp instructions" + (f" = {tot / units:.2f} per unit ({units:.0f} units)" if units else ""))
    for k, n in by.most_common(18):
        per = f"{n / units:7.2f}/unit" if units else f"{100 * n / tot:5.1f}%"
        print(f"  {k:10s} {per}   stall samples {100 * samp[k] / ts:5.1f}%")
    print("SASS lines with the most stall samples:")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:12]
    for i in sorted(top):
        r = data[i]
        print(f"  {100 * int(r[isamp]) / ts:5.2f}%  x{int(r[ie]):>9d}  {r[ia].strip()[:90]}")


if __name__ == "__main__":
    main()
