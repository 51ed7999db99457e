"""Hot spots of one kernel from an ncu report: per source line (needs -lineinfo) or per SASS instruction.
usage: ncu_hot.py <report.ncu-rep> <kernel regex> [sass|line] [top N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "line"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
view = "source,sass" if mode == "line" else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass" if mode == "line" else "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line"))
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
samp = ix.get("Warp Stall Sampling (All Samples)") or ix.get("# Samples")
inst = ix.get("Instructions Executed")
tot_s = tot_i = 0
items = []
for r in rows[hdr_i + 1:]:
    if len(r) <= max(samp, inst):
        continue
    try:
        s = int(r[samp] or 0)
        n = int(r[inst] or 0)
    except ValueError:
        continue
    if any(a == r[0] for _, _, _, a in items[-400:]):
        continue
    items.append((s, n, r[ix.get("Source", 1)][:150], r[0]))
tot_s = sum(i[0] for i in items)
tot_i = sum(i[1] for i in items)
print(f"total samples {tot_s}, instructions executed {tot_i}")
for s, n, src, a in sorted(items, reverse=True)[:top]:
    print(f"{100.0 * s / max(tot_s, 1):6.2f}%  inst {100.0 * n / max(tot_i, 1):6.2f}%  {a[-5:]}  {src}")
