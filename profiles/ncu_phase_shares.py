"""Stall samples of a kernel split at its BAR.SYNC instructions (needs ncu on PATH).
usage: ncu_phase_shares.py report.ncu-rep kernel-regex [out.txt]

Samples at the first instructions after a barrier are warps waiting AT that barrier (the sampler
reports the next instruction to issue), i.e. time the CTA spends waiting for the slowest warp of
the phase before it."""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[starts[0]]
body = [r for r in rows[starts[0] + 1:(starts[1] - 1 if len(starts) > 1 else len(rows))] if len(r) == len(hdr)]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
total = sum(int(r[si] or 0) for r in body)
lines = [f"kernel {rx}: {len(body)} SASS instructions, {total} stall samples (first launch in the report)",
         "phase  last_instr  samples  share  warp_instructions  samples_in_first_8_instrs(=waiting at the previous barrier)"]
phase, acc, ex, first8, since = 0, 0, 0, 0, 0
for k, r in enumerate(body):
    n = int(r[si] or 0)
    acc += n
    ex += int(r[ie] or 0)
    if since < 8:
        first8 += n
    since += 1
    if "BAR.SYNC" in r[so] or k == len(body) - 1:
        lines.append(f"{phase:5d}  {k:10d}  {acc:7d}  {acc / max(total, 1):5.1%}  {ex:17d}  {first8}")
        phase, acc, ex, first8, since = phase + 1, 0, 0, 0, 0
top = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:12]
lines.append("top instructions by samples:")
for k in top:
    lines.append(f"  {body[k][si]:>6}  #{k:<6} {body[k][so].strip()}")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(out + "\n")
