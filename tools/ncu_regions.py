#!/usr/bin/env python
"""Stall samples of one captured launch binned along the instruction stream (usage: ncu_regions.py REP [LAUNCH] [BIN])."""
import csv, io, subprocess, sys
rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r)
hdr = rows[hi]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall = [(k, i) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
data, seen = [], set()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ia] in seen:
        continue
    seen.add(r[ia])
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    data.append((r[isrc], s, {k: int(r[i] or 0) for k, i in stall}))
tot = sum(d[1] for d in data) or 1
print("instructions", len(data), "samples", tot)
for b in range(0, len(data), B):
    chunk = data[b:b + B]
    s = sum(d[1] for d in chunk)
    if 100.0 * s / tot < 0.7:
        continue
    ops, st = {}, {}
    for d in chunk:
        t = d[0].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
        for k, v in d[2].items():
            st[k] = st.get(k, 0) + v
    top = sorted(st.items(), key=lambda t: -t[1])[:3]
    topops = sorted(ops.items(), key=lambda t: -t[1])[:4]
    hot = max(chunk, key=lambda d: d[1])
    print("%5d-%5d %5.1f%% %-58s %-46s hot: %s (%.1f%%)" % (
        b, b + B, 100.0 * s / tot, str(topops), str([(k[6:], round(100 * v / max(s, 1))) for k, v in top]),
        hot[0][:40], 100.0 * hot[1] / tot))
