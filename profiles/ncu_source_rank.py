"""Rank CUDA source lines of a kernel by warp-stall samples (or executed instructions) from an
.ncu-rep.  usage: python profiles/ncu_source_rank.py report.ncu-rep [top_n] [kernel-substr] [inst]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ksub = sys.argv[3] if len(sys.argv) > 3 else ""
by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None
kernel = None
aggs = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kernel = r[1]
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("Warp Stall Sampling (All Samples)")
        ii = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0].strip():
        try:
            s = int(r[si] or 0)
            n = int(r[ii] or 0) if ii is not None else 0
        except ValueError:
            continue
        key = (fname, r[0], r[1].strip()[:100])
        a = aggs.setdefault(kernel, {}).setdefault(key, [0, 0])
        a[0] += s
        a[1] += n
seen = set()
for kernel, agg in aggs.items():
    if ksub not in (kernel or ""):
        continue
    short = (kernel or "?")[:60]
    if short in seen:
        continue
    seen.add(short)
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"=== {short}: samples {tot}, warp-instructions {toti}")
    order = sorted(agg.items(), key=lambda kv: -(kv[1][1] if by_inst else kv[1][0]))[:top]
    for k, v in order:
        print(f"{v[0]:6d} {100 * v[0] / tot:5.1f}%  inst={v[1]:9d} {100 * v[1] / toti:5.1f}%  {k[0]}:{k[1]}: {k[2]}")
