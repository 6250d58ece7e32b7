#!/usr/bin/env python
"""Summarise an ncu --page source --print-source cuda,sass --csv dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
ix = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
wf = hdr.index('L1 Wavefronts Shared'); wfi = hdr.index('L1 Wavefronts Shared Ideal')
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        lines.append((int(r[0]), r[1].strip(), int(r[ix]), int(r[isamp]), int(r[wf]), int(r[wfi])))
    except ValueError:
        pass
tot = sum(l[2] for l in lines); ts = sum(l[3] for l in lines); tw = sum(l[4] for l in lines); twi = sum(l[5] for l in lines)
print(f"total warp-inst {tot}  samples {ts}  smem wavefronts {tw} (ideal {twi})")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"L{l[0]:<4} inst {100*l[2]/tot:5.1f}%  samp {100*l[3]/ts:5.1f}%  wf {l[4]:>9} ideal {l[5]:>9} | {l[1][:100]}")
