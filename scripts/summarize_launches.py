"""Summarise an `ncu --metrics gpu__time_duration.sum[,...] --csv` launch list per kernel (durations only)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print(f"total {s:.0f} us over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.1f} us {100 * v / s:5.1f}%  n={cnt[k]:4d}  avg={v / cnt[k]:8.1f}  {k[:90]}")
