"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`): launches, total us, share.
usage: python tools/summarise_launches.py gpurun_out/launches_bench.csv > profiles/roundN_launches_bench_summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    name = re.sub(r"\(.*", "", r[kn])[:70]
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {total:.1f} us total (ncu serialises launches and runs them cold: shares, not absolutes)")
print("launches   us total   share  kernel")
for name, us in tot.most_common():
    print(f"{cnt[name]:8d} {us:10.1f} {100 * us / total:6.1f}%  {name}")
