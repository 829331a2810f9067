#!/usr/bin/env python
"""Per-CUDA-line executed-instruction counts from `ncu -i X.ncu-rep --page source --print-source cuda --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = None
tot = 0
out = []
fname = ""
for r in rows:
    if 'Source' in r and 'Instructions Executed' in r:
        hdr = r; si = hdr.index('Source'); ii = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) <= ii:
        if r and r[0].startswith('File'): fname = r[1] if len(r) > 1 else r[0]
        continue
    try: n = int(r[ii])
    except ValueError: continue
    tot += n
    out.append((n, fname.split('/')[-1], r[0], r[si][:120]))
print("total warp instructions", tot)
for n, f, l, s in out:
    if n > tot * thr: print("%9d %5.1f%% %s:%s: %s" % (n, 100 * n / tot, f, l, s))
