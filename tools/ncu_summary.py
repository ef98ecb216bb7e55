#!/usr/bin/env python
"""Summarise one launch of an .ncu-rep (usage: ncu_summary.py REP [TOP_N] [LAUNCH_INDEX]): headline metrics, stall reasons, hottest SASS lines."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # which captured launch (0-based)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + launch]
def g(name):
    for i, k in enumerate(hdr):
        if k == name: return vals[i] + " " + units[i]
    return "n/a"
for k in ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads", "smsp__inst_executed.sum",
          "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
    print("%-70s %s" % (k, g(k)))
st = {}
for i, k in enumerate(hdr):
    if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
        try: st[k[len("smsp__pcsamp_warps_issue_stalled_"):]] = int(vals[i])
        except ValueError: pass
tot = sum(st.values()) or 1
print("stall reasons (pc samples): " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(st.items(), key=lambda t: -t[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, isamp, iexe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(k, i) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
seen, data, byop, byexe = set(), [], collections.Counter(), collections.Counter()
ia = hdr.index("Address")
for r in rows[2:]:
    if len(r) < len(hdr) or r[ia] in seen: continue
    seen.add(r[ia])
    try: s = int(r[isamp])
    except ValueError: continue
    data.append((s, r))
    toks = r[isrc].split()
    op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
    byop[op] += s
    try: byexe[op] += int(r[iexe])
    except ValueError: pass
tot = sum(s for s, _ in data) or 1
print("samples by opcode: " + ", ".join("%s %.1f%% (%.2e exec)" % (op, 100.0 * s / tot, byexe[op]) for op, s in byop.most_common(12)))
print("hottest instructions:")
for s, r in sorted(data, key=lambda t: -t[0])[:top_n]:
    stl = sorted(((k, int(r[i] or 0)) for k, i in stall_cols), key=lambda t: -t[1])[:2]
    print("  %5.2f%%  %-60s %s" % (100.0 * s / tot, r[isrc][:60], stl))
