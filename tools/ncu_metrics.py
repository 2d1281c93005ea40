"""Extract the metrics quoted in DESIGN.md / profiles/ from an ncu report:
   ncu -i X.ncu-rep --page raw --csv | python tools/ncu_metrics.py <out.txt> "<one-line description>" """
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
h, u, v = rows[0], rows[1], rows[2]
pat = re.compile(r"^(dram__bytes_(read|write)\.sum(\.per_second)?$|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate\.pct"
                 r"|l1tex__t_sectors_pipe_lsu_mem_(global|local)_op_(ld|st|atom|red)\.sum$|launch__(block_size|grid_size|registers_per_thread|occupancy_limit_\w+|shared_mem_per_block_static)"
                 r"|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum$|lts__throughput\.avg\.pct"
                 r"|smsp__pcsamp_warps_issue_stalled_\w+$|smsp__inst_executed_op_(local|shared|global)\w*\.sum$|smsp__average_warps_issue_stalled_\w+_per_issue_active)")
with open(sys.argv[1], "w") as out:
    out.write(sys.argv[2] + "\n")
    for n, un, val in sorted(zip(h, u, v)):
        if pat.search(n):
            out.write("%-96s %-12s %s\n" % (n, un, val))
