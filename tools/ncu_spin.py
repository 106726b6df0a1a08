#!/usr/bin/env python
"""Summarises an `ncu --import-source on` capture of the rollout kernel: executed instructions per pair-step for the
mover loop, the observer loop and everything else, how often the polling loops spin, and the top stall reasons.
Usage: python tools/ncu_spin.py report.ncu-rep [n_pairs n_steps]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ia, isrc, iex, ism = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = int(data[0][ia], 16)
ins = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[iex]), int(r[ism])) for r in data]
ps = pairs * steps
tot = sum(x[2] for x in ins); stot = sum(x[3] for x in ins)
# loops = backward branches whose body executes ~once per pair-step
loops = []
for a, s, e, sm in ins:
    m = re.search(r'BRA\s+0x([0-9a-f]+)', s)
    if m:
        tgt = int(m.group(1), 16) - base
        if 0 <= tgt < a and a - tgt > 150 * 16 and e > 0.5 * ps:
            loops.append((tgt, a))
print('total executed %.1f instr/pair-step (%d), samples %d' % (tot / ps, tot, stot))
covered = 0
for lo, hi in loops:
    ex = sum(x[2] for x in ins if lo <= x[0] <= hi); sm = sum(x[3] for x in ins if lo <= x[0] <= hi)
    spin = [(x[0], x[1], x[2] / ps) for x in ins if lo <= x[0] <= hi and x[2] > 1.2 * ps and ('LDS' in x[1])]
    print('loop %05x..%05x: %d static, %.1f executed/pair-step, %.1f%% of samples; polls/step: %s'
          % (lo, hi, (hi - lo) // 16 + 1, ex / ps, 100.0 * sm / stot, ', '.join('%.2f' % s[2] for s in spin) or '-'))
    covered += ex
print('outside the loops: %.1f instr/pair-step' % ((tot - covered) / ps))
