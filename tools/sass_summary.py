#!/usr/bin/env python
"""Static SASS summary of the rollout kernel the bench times: opcode histogram of the whole function and of its two
step loops (mover / observer), plus the memory-movement mnemonics that tell an
Ampere-style kernel from an sm_100 one (LDGSTS = cp.async, UBLKCP / UTMA* = bulk / tensor TMA, SYNCS = mbarrier).
No GPU needed (cuobjdump on the in-tree library).   Usage: python tools/sass_summary.py [kernel regex] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'atc_reinforcement_learning_b200', 'csrc', 'libatc_b200.so')
WATCH = ['LDGSTS', 'LDGDEPBAR', 'DEPBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'UTMACMDFLUSH', 'SYNCS', 'UTCHMMA', 'UTCQMMA',
         'LDS', 'STS', 'LDG', 'STG', 'LDC', 'LDCU', 'SHFL', 'VOTE', 'MUFU', 'DFMA', 'DADD', 'DMUL', 'DSETP', 'FFMA',
         'FFMA2', 'FMUL2', 'FADD2', 'F2F', 'F2I', 'I2F', 'BAR', 'MEMBAR', 'WARPSYNC', 'BRA', 'CALL', 'RET', 'ACQBULK',
         'NANOSLEEP', 'CCTL']


def functions():
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    cur, out = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(anonymous namespace\)::', '', cur)
            out[cur] = []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and cur is not None:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    ins = re.sub(r'^@!?U?P\d+\s+', '', ins)
    return ins.split()[0].split('.')[0] if ins else '?'


def hist(rows):
    return collections.Counter(opcode(i) for _, i in rows)


def loops(rows):
    """(target, branch address) of every backward branch, widest first"""
    res = []
    for a, i in rows:
        m = re.search(r'\bBRA\b.*?(0x[0-9a-f]+)', i)
        if m:
            t = int(m.group(1), 16)
            if t < a:
                res.append((a - t, t, a))
    return sorted(res, reverse=True)


def row_store_spans(rows):
    """Instruction distance between the first and the last of the five 8-byte streaming stores of each observation row
    (STG.E.EF.64 with the same base register, offsets 0 .. 0x20) — 4 when they are issued back to back.  The L1 merges
    the partial-sector writes of adjacent stores only; a scheduler that spreads them costs the kernel 12 %."""
    groups = collections.OrderedDict()
    for idx, (a, ins) in enumerate(rows):
        m = re.search(r'STG\.E\.EF\.64 desc\[\w+\]\[(R\d+)\.64(?:\+(0x[0-9a-f]+))?\]', ins)
        if m:
            off = int(m.group(2), 16) if m.group(2) else 0
            key = m.group(1)
            g = groups.setdefault(key, [])
            if g and (idx - g[-1][0] > 64 or any(o == off for _, o in g)):     # a new row through the same register
                key = '%s@%d' % (key, idx)
                g = groups.setdefault(key, [])
            g.append((idx, off))
    return [max(i for i, _ in g) - min(i for i, _ in g) for g in groups.values()
            if sorted(o for _, o in g) == [0, 8, 16, 24, 32]]


def show(title, rows):
    h = hist(rows)
    print('%s: %d instructions' % (title, len(rows)))
    print('   watched : ' + '  '.join('%s %d' % (k, h[k]) for k in WATCH if h.get(k)))
    print('   top     : ' + '  '.join('%s %d' % kv for kv in h.most_common(24)))


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else r'atc_rollout_pipe_kernel<4, false, false, 2, 14>'
    fns = functions()
    print('# SASS summary of %s (cuobjdump -sass, sm_100a); kernels in the library: %d' % (os.path.relpath(LIB, ROOT), len(fns)))
    tot = collections.Counter()
    for rows in fns.values():
        tot.update(hist(rows))
    print('whole library, movement / sync mnemonics: ' + '  '.join('%s %d' % (k, tot[k]) for k in
          ['LDGSTS', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'UTCHMMA', 'FFMA2', 'FMUL2', 'FADD2'] ))
    for name, rows in fns.items():
        if not re.search(pat, name):
            continue
        print()
        show(name, rows)
        # the step loops, by what they contain: the mover publishes its message with shared-memory stores and writes
        # nothing to global memory; the observer holds the observation stores (STG) — widest loop of each kind
        picked = []
        for role, want in (('mover', lambda h: h['STS'] >= 5 and h['STG'] == 0 and h['DFMA'] >= 10),
                           ('observer', lambda h: h['STG'] >= 10)):
            for span, t, a in loops(rows):
                if want(hist([r for r in rows if t <= r[0] <= a])):
                    picked.append((role, t, a))
                    break
        for role, t, a in picked:
            sel = [r for r in rows if t <= r[0] <= a]
            show('   %s step loop [%05x, %05x]' % (role, t, a), sel)
            if role == 'observer':
                print('   row stores (5 x STG.E.EF.64 per 40-byte row): instruction span per row %s (4 = back to back)'
                      % row_store_spans(sel))
        if picked:
            rest = [r for r in rows if not any(t <= r[0] <= a for _, t, a in picked)]
            show('   outside the step loops (prologue, epilogue, out-of-line paths)', rest)


if __name__ == '__main__':
    main()
