"""Opcode mix per pixel from an `ncu --page source --csv` dump.
usage: python profiles/opmix.py dump.csv n_pixels [kernel-substring]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2])
pick = sys.argv[3] if len(sys.argv) > 3 else None
cur = None
hdr = None
ops = collections.Counter()
samp = collections.Counter()
tot = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        if tot and cur:
            break
        cur = r[1] if (pick is None or pick in r[1]) else None
        continue
    if r and r[0] == "Address":
        hdr = r
        iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if cur is None or hdr is None or len(r) <= iE:
        continue
    src = r[iA].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:12]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "F2I", "I2F")) else op.split(".")[0]
    n = int(r[iE])
    tot += n
    ops[op] += n
    samp[op] += int(r[iS])
print(cur)
print("thread-instructions per pixel: %.1f" % (tot * 32 / npx))
for op, n in ops.most_common(40):
    print(f"{op:14s} {n * 32 / npx:7.2f}/px   stall samples {samp[op]}")
