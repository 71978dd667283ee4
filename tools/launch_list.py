"""Print the last step's kernels from an ncu gpu__time_duration launch list (csv)."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hdr = rows[0]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
data = [(r[ki], r[gi], float(r[vi].replace(",", "")) / (1000 if r[ui].startswith("n") else 1)) for r in rows[1:]]
n = len(data) // steps
tot = sum(v for _, _, v in data[-n:])
for k, g, v in data[-n:]:
    print(f"{k[:58]:58s} {g:16s} {v:9.2f} us {100 * v / tot:5.1f}%")
print(f"total {tot:.1f} us over {n} launches")
