#!/usr/bin/env python
"""Per-role stall breakdown of the pipelined rollout kernel from `ncu --page source --csv` / `--page raw --csv` dumps:
finds the two step loops (backward branches), prints instructions per iteration, stall-reason shares and the
instructions the warps wait at."""
import csv
import re
import sys


def main(src_csv, raw_csv, n_iters):
    rows = list(csv.reader(open(raw_csv)))
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if any(w in h for w in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
                                'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread',
                                'launch__occupancy_limit', 'dram__bytes_read.sum', 'dram__bytes_write.sum']) \
                and 'per_second' not in h and 'pct_of' not in h.replace('issue_active.avg.pct_of_peak_sustained_active', ''):
            print(h, u, v)
    rows = list(csv.reader(open(src_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[hi + 1:]
    base = int(data[0][ix['Address']], 16)
    addr = lambda r: int(r[ix['Address']], 16) - base
    cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot_all = sum(float(r[ix['# Samples']] or 0) for r in data)
    loops = []
    for r in data:
        m = re.search(r'BRA (0x[0-9a-f]+)', r[ix['Source']])
        if m and float(r[ix['Instructions Executed']] or 0) / n_iters > 0.9:
            tgt = int(m.group(1), 16) - base
            if tgt < addr(r) - 0x400:
                loops.append((tgt, addr(r) + 0x10))
    bounds = loops + [(loops[-1][1] if loops else 0, 1 << 30)]
    for k, (lo, hi_) in enumerate(bounds):
        name = ['mover loop', 'observer loop', 'out of line'][min(k, 2)]
        sel = [r for r in data if lo <= addr(r) < hi_]
        tot = sum(float(r[ix['# Samples']] or 0) for r in sel)
        ninst = sum(float(r[ix['Instructions Executed']] or 0) for r in sel) / n_iters
        print('%s [%05x, %05x): samples %.1f%%, instructions/iteration %.1f' % (name, lo, hi_, 100 * tot / tot_all, ninst))
        for c in cols:
            t = sum(float(r[ix[c]] or 0) for r in sel)
            if tot and t / tot > 0.015:
                print('   %-26s %5.1f%%' % (c, 100 * t / tot))
        for r in sorted(sel, key=lambda r: -float(r[ix['# Samples']] or 0))[:8]:
            print('      %05x %5.2f%% x%.2f %s' % (addr(r), 100 * float(r[ix['# Samples']]) / tot_all,
                                                  float(r[ix['Instructions Executed']] or 0) / n_iters,
                                                  ' '.join(r[ix['Source']].split())[:70]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]))
