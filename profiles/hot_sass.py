"""Print the hottest SASS instructions (by executed warp instructions) of an ncu report's source page."""
import csv, subprocess, sys
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(txt[1:]))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
ex, src, smp = ci["Instructions Executed"], ci["Source"], ci["# Samples"]
data = [(int(r[ex]), int(r[smp]), r[src].strip()) for r in rows[1:] if len(r) > ex and r[ex].isdigit()]
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"total warp instructions {tot}, samples {ts}, distinct SASS lines {len(data)}")
cnt = {}
for n, s, t in data:
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    cnt[op.split(".")[0]] = cnt.get(op.split(".")[0], 0) + n
print("by opcode:", ", ".join(f"{k}:{100*v/tot:.1f}%" for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:14]))
for n, s, t in data:
    if n > tot * frac or s > ts * frac * 2:
        print(f"{100*n/tot:5.1f}% inst {100*s/max(ts,1):5.1f}% samples  {t[:100]}")
