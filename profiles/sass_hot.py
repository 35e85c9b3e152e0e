"""Top SASS instructions by warp-stall samples with context:  python profiles/sass_hot.py report.ncu-rep [top_n] [context]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > hdr.index("# Samples")]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[iS] or 0) for r in data)
print("total samples", tot)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iS] or 0))[:top]
for i in sorted(idx):
    print(f"--- {100.0 * int(data[i][iS]) / tot:5.1f}%  sass line {i}")
    for k in range(max(0, i - ctx), i + 1):
        print("      ", data[k][iSrc][:120])
