"""Launch list (kernel, launches, total / mean time, share) from an
`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` capture.  usage: launch_list.py X.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
order = []
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
    name = re.sub(r"\(.*$", "", r[ik]).replace("void vcfb::<", "").replace("vcfb::", "")
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += us
    order.append((name.split("<")[0], us))
total = sum(d[1] for d in tot.values())
print("launches     total us   mean us   share  kernel")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:8d} {t:12.1f} {t / n:9.1f} {100 * t / total:6.2f}%  {name}")
print("# first launches in order:", ", ".join(f"{n} {int(u)}" for n, u in order[:16]))
