"""Per-instruction execution counts and stall samples of a kernel in an ncu report (needs --import-source on).
    python tools/ncu_sass.py report.ncu-rep [min_count]"""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if 'Source' in r)
hdr = rows[h]
ie, src, smp, thr = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Avg. Threads Executed')
kernel = sys.argv[3] if len(sys.argv) > 3 else None   # substring of the kernel name when the report holds several
data, on = [], kernel is None
for r in rows[h + 1:]:
    if r and r[0] == 'Kernel Name':
        on = kernel is None or kernel in r[1]
        if on:
            data = []   # a kernel can appear once per view: keep the last one
        continue
    if on and len(r) > ie and r[ie] != 'Instructions Executed':
        data.append(r)
tot = sum(int(r[ie] or 0) for r in data)
tots = sum(int(r[smp] or 0) for r in data)
print('total warp instructions', tot, 'samples', tots)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in data:
    n = int(r[ie] or 0)
    if n >= lo:
        print(f"{n:10d} {100.0 * n / tot:5.1f}% smp {100.0 * int(r[smp] or 0) / max(tots, 1):5.1f}% thr {r[thr]:>5s}  {r[src][:100]}")
