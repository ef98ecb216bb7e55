#!/usr/bin/env python
"""L1 / LSU metrics of one launch of an .ncu-rep (usage: ncu_l1.py REP [LAUNCH_INDEX]) - the numbers that exposed the ZGP kernel's limit."""
import csv, io, subprocess, sys
rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + launch]
for k in ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
          "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    if k in hdr:
        i = hdr.index(k)
        print("%-84s %s %s" % (k, vals[i], units[i]))
