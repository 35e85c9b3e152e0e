"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"^void |msam::", "", name)
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:8.3f} ms {100*v/T:5.1f}%  n={cnt[k]:4d}  avg {1e3*v/cnt[k]:8.1f} us  {k}")
