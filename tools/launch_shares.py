"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total ms, share.
usage: python tools/launch_shares.py launches.csv [header text]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[hdr + 1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = r[ik].split("(")[0].split("<")[0].replace("rvgp::", "").replace("void ", "")
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# total kernel ms = %.1f over %d launches (per-launch times are cold-cache / serialised: compare SHARES)" % (total, sum(cnt.values())))
for name in sorted(tot, key=lambda n: -tot[n]):
    print("%-44s launches %6d  ms %9.2f  share %5.1f%%  avg us %8.1f" % (name[:44], cnt[name], tot[name], 100 * tot[name] / total, 1e3 * tot[name] / cnt[name]))
