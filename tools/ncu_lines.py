"""Per-source-line instruction counts of one kernel: joins the SASS page of an .ncu-rep with the
line table nvdisasm prints for the cubin.  python tools/ncu_lines.py rep kernel_substr cubin mangled_substr"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
kfilter = ["-k", sys.argv[4]] if len(sys.argv) > 4 else []
by_samples = len(sys.argv) > 5
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# instruction index -> source line, for the function
in_fn, line_of, cur = False, [], None
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        in_fn = mangled in ln
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        line_of.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + kfilter, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [r for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
print(f"sass rows {len(sass)}  disasm instrs {len(line_of)}")
agg, samp = defaultdict(int), defaultdict(int)
for i, r in enumerate(sass):
    key = line_of[i] if i < len(line_of) else None
    agg[key] += int(r[ie]); samp[key] += int(r[ss])
tot = sum(agg.values()); ts = sum(samp.values())
print(f"total warp-instructions {tot}  samples {ts}")
for key, v in sorted(agg.items(), key=lambda kv: (-samp[kv[0]] if by_samples else -kv[1]))[:45]:
    print(f"{str(key):32s} inst {v:12d} {100*v/tot:5.1f}%   samples {100*samp[key]/max(ts,1):5.1f}%")
