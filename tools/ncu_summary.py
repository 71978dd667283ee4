"""Summarise an .ncu-rep raw page: python tools/ncu_summary.py report.ncu-rep [regex]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|lts__t_bytes.sum$|"
                 r"lts__throughput.avg.pct|l1tex__throughput.avg.pct|sm__throughput.avg.pct|smsp__issue_active.avg.pct|"
                 r"sm__warps_active.avg.pct|launch__(registers_per_thread|occupancy_limit|grid_size|block_size|shared_mem)|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared|l1tex__data_pipe_lsu_wavefronts(_mem_shared)?.sum$|"
                 r"warp_issue_stalled.*per_warp_active.pct|sm__inst_executed_pipe_(lsu|fma|alu|fmaheavy|xu).sum$|"
                 r"smsp__inst_executed.sum$|sm__cycles_elapsed.max|sm__cycles_active.avg|l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$|"
                 r"l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$|lts__t_sectors_srcunit_tex_op_read.sum$|"
                 r"l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|sm__pipe_.*cycles_active.avg.pct")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f"  {h:95s} {units[i]:14s} {r[i]}")
