#!/usr/bin/env python
"""Executed-instruction histogram by opcode from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
hist = collections.Counter(); tot = 0; thr = 0
listing = []
for r in rows:
    if 'Source' in r and 'Instructions Executed' in r:
        hdr = r; si = r.index('Source'); ii = r.index('Instructions Executed'); ti = r.index('Thread Instructions Executed'); continue
    if hdr is None or len(r) <= ti: continue
    try: n = int(r[ii]); t = int(r[ti])
    except ValueError: continue
    s = r[si].strip()
    parts = s.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0].rstrip(';')
    hist[op] += n; tot += n; thr += t
    listing.append((n, t, s))
print("total warp instr", tot, " avg active threads %.1f" % (thr / max(tot, 1)))
for op, n in hist.most_common(40): print("%10d %5.1f%% %s" % (n, 100.0 * n / tot, op))
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
        for n, t, s in listing: f.write("%9d %5.1f %s\n" % (n, t / max(n, 1), s))
