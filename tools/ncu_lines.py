#!/usr/bin/env python
"""Per-instruction stall samples of one captured launch (usage: ncu_lines.py REP FIRST LAST [LAUNCH]):
instruction index, share of all samples, top two stall reasons, SASS."""
import csv, io, subprocess, sys
rep, first, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
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
for i in range(first, min(last, len(data))):
    src_, s, st = data[i]
    top = sorted(st.items(), key=lambda t: -t[1])[:2]
    print("%5d %5.2f%% %-34s %s" % (i, 100.0 * s / tot, " ".join("%s:%d" % (k[6:], v) for k, v in top if v), src_[:70]))
