#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list as a markdown table."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = {}
for r in rows:
    name = re.sub(r"\(.*$", "", r[4].replace("void ", "").replace("unnamed>::", "").replace("mclcar::<", ""))
    k = (name, r[7], r[8])
    a = agg.setdefault(k, [0, 0])
    a[0] += 1
    a[1] += int(float(r[14]))
tot = sum(a[1] for a in agg.values())
print("| kernel | block | grid | launches | total us | avg us | share |")
print("|---|---|---|---:|---:|---:|---:|")
for (name, blk, grd), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {blk} | {grd} | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.1f} | {100.0 * ns / tot:.1f}% |")
print(f"\ntotal {tot / 1e6:.3f} ms in {len(rows)} launches")
