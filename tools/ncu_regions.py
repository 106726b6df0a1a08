#!/usr/bin/env python
"""Per-region instruction / stall-sample totals from an `ncu --page source --csv` dump (see profiles/README.md)."""
import csv
import re
import sys


def load(path, n_cta_steps):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    data = rows[hi + 1:]
    tot = sum(float(r[ix['# Samples']] or 0) for r in data)
    base = int(data[0][ix['Address']], 16)
    out = []
    for r in data:
        out.append((int(r[ix['Address']], 16) - base, float(r[ix['Instructions Executed']] or 0) / n_cta_steps,
                    100 * float(r[ix['# Samples']] or 0) / tot, r[ix['Source']].strip()))
    return out


if __name__ == '__main__':
    data = load(sys.argv[1], float(sys.argv[2]))
    for a, ie, sp, s in data:
        print('%05x %6.3f %5.2f %s' % (a, ie, sp, re.sub(r'\s+', ' ', s)))
