"""Aggregate ncu warp-stall samples per CUDA source line:  python profiles/line_samples.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.Counter()
text = {}
fname = "?"
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if "# Samples" in r:
        hdr = r
        iS, iL, iT = r.index("# Samples"), r.index("Line No"), r.index("Source")
        continue
    if hdr is None or len(r) <= iS:
        continue
    try:
        n = int(r[iS] or 0)
    except ValueError:
        continue
    key = (fname, r[iL])
    agg[key] += n
    if r[iT].strip():
        text.setdefault(key, r[iT].strip())
tot = sum(agg.values())
print("total samples", tot)
for (f, l), n in agg.most_common(top):
    print(f"{100.0 * n / tot:5.1f}%  {f}:{l}  {text.get((f, l), '')[:110]}")
