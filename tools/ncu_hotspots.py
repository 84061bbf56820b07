#!/usr/bin/env python
"""Per-phase stall summary from `ncu -i X.ncu-rep --page source --csv` (SASS view): splits every kernel at its BAR.SYNC
instructions and prints, per segment, the share of warp-stall samples and the dominant stall reasons; then the hottest
single instructions.   python tools/ncu_hotspots.py gpurun_out/X.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

path = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', out)
seen = set()
for blk in blocks[1:]:
    lines = blk.split("\n")
    name = lines[0].strip().strip('",')
    if name in seen or (pat and not pat.search(name)):
        continue
    seen.add(name)
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    total = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    print(f"\n## {name}\n   {len(data)} SASS instructions, {total} stall samples")
    seg, segs = [], []
    for r in data:
        seg.append(r)
        if "BAR.SYNC" in r[ix["Source"]]:
            segs.append(seg)
            seg = []
    segs.append(seg)
    for k, sg in enumerate(segs):
        n = sum(int(r[ix["# Samples"]] or 0) for r in sg)
        reasons = {c: sum(int(r[ix[c]] or 0) for r in sg) for c in stall_cols}
        top = sorted(reasons.items(), key=lambda kv: -kv[1])[:4]
        ninst = len(sg)
        ops = {}
        for r in sg:
            op = r[ix["Source"]].split()[0].split(".")[0] if r[ix["Source"]].split() else "?"
            if op.startswith("@"):
                op = r[ix["Source"]].split()[1].split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        topops = ", ".join(f"{o} {c}" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"   phase {k}: {ninst:5d} instr, {100.0 * n / total:5.1f} % of samples | "
              + ", ".join(f"{c[6:]} {100.0 * v / max(n, 1):.0f}%" for c, v in top) + f" | {topops}")
    hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]
    print("   hottest instructions:")
    for r in hot:
        n = int(r[ix["# Samples"]] or 0)
        reasons = sorted(((c, int(r[ix[c]] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"     {100.0 * n / total:4.1f} %  {r[ix['Address']][-5:]}  {r[ix['Source']][:70]:70s} "
              + ", ".join(f"{c[6:]} {v}" for c, v in reasons))
