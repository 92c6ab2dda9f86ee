"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel name: count, total, share of the step."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mi = hdr.index("Metric Name") if "Metric Name" in hdr else None
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if mi is not None and r[mi] != "gpu__time_duration.sum":
        continue                      # captures with several metrics per launch: durations only
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"msmc::<unnamed>::|at::native::|<unnamed>::", "", name)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print("launches %d, sum of kernel durations %.2f ms" % (sum(cnt.values()), total / 1000.0))
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%9.3f ms %5.1f%% %6d x %8.1f us  %s" % (v / 1000.0, 100.0 * v / total, cnt[k], v / cnt[k], k[:110]))
