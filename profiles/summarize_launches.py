#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python profiles/summarize_launches.py profiles/r1_launches_v4.csv > profiles/r1_launches_v4.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = rows[hdr]
kn, mv, mu = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
tot = collections.Counter()
cnt = collections.Counter()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1e-6)
    except ValueError:
        continue
    name = r[kn].split("(")[0]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print("# (configs[1], batch 1e6; cold-cache serialised launches: compare SHARES, not absolutes)")
print(f"# total {total:.3f} ms over {sum(cnt.values())} launches (set-up + fp64 peak probe + 3 warm-up + 2 timed steps + host-path steps)")
for name, v in tot.most_common():
    print(f"{v:12.3f} ms {100 * v / total:6.2f} %  x{cnt[name]:<4d} {name[:100]}")
