"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.OrderedDict(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:78]
    tot[name] = tot.get(name, 0) + v
    cnt[name] += 1
s = sum(tot.values())
print(f"total {s:.1f} us over {sum(cnt.values())} launches (cold-cache, serialised: compare shares)")
mine = sum(v for k, v in tot.items() if "sdetr::" in k)
print(f"sdetr:: kernels {mine:.1f} us ({100 * mine / s:.1f}%), {sum(c for k, c in cnt.items() if 'sdetr::' in k)} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v:9.1f} us {100 * v / s:5.1f}% x{cnt[k]:3d}  {k}")
