#!/usr/bin/env python
"""Histogram of executed SASS opcodes (and stall samples) from an exported `--page source --csv --print-source sass`
file:  python profiles/sass_hist.py gpurun_out/r1c_k6.sass.csv.gz [top]"""
import collections, csv, gzip, sys
fh = gzip.open(sys.argv[1], "rt")
r = csv.reader(fh)
next(r)
hdr = next(r)
i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
i_wf = hdr.index("L1 Wavefronts Shared")
ops = collections.Counter(); smp = collections.Counter(); wf = collections.Counter()
tot = 0
for row in r:
    if len(row) <= i_ex: continue
    toks = row[i_src].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "LDC", "SHFL", "LDGSTS")) and "." in op else "")
    n = int(row[i_ex] or 0)
    ops[op] += n; tot += n
    smp[op] += int(row[i_smp] or 0)
    wf[op] += int(row[i_wf] or 0)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ts = sum(smp.values())
print(f"total warp instructions {tot}, samples {ts}")
for op, n in ops.most_common(top):
    print(f"{op:14s} {n:12d} {100*n/tot:6.2f}%  samples {100*smp[op]/max(ts,1):6.2f}%  smem wavefronts {wf[op]}")
