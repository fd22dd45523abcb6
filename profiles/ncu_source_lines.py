"""Executed warp instructions and stall samples per CUDA source line of one kernel (needs -lineinfo and
--import-source on).  usage: ncu_source_lines.py report.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
fname, agg = None, {}
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No" or len(r) < 8 or r[2] != "-":
        continue
    try:
        agg[(fname, int(r[0]), r[1].strip()[:100])] = (int(r[6] or 0), int(r[7] or 0))
    except ValueError:
        pass
ts, ti = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print(f"{ts} stall samples, {ti} warp instructions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100 * v[1] / ti, 100 * v[0] / ts, k[0], k[1], k[2]))
