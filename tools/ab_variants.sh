#!/bin/bash
# A/B timing of kernel variants: every variants/*.so is copied over the in-tree library and benched (device-resident
# timing; the in-run parity check against the oracle on the first repetition).  Repetitions are interleaved, the
# summary prints the median and the best of each variant.  Usage (on the GPU box): REPS=3 bash tools/ab_variants.sh [bench args]
LIB=atc_reinforcement_learning_b200/csrc/libatc_b200.so
REPS=${REPS:-3}
cp $LIB /tmp/orig.so
: > /tmp/ab_all.txt
for rep in $(seq 1 $REPS); do
for v in variants/*.so; do
  cp $v $LIB
  EXTRA="--skip-parity"; [ $rep = 1 ] && EXTRA=""
  ENVF=${v%.so}.env; ENVS=""; [ -f $ENVF ] && ENVS=$(cat $ENVF)      # optional environment of the variant (one line)
  env $ENVS python bench.py --steps 20 --warmup 5 --skip-extras $EXTRA "$@" 2>/tmp/ab.err | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); p=d.get('parity_in_run') or {}
    print('$v', round(d['value']/1e9,3), round(d['roofline']['frac'],4), 'parity', p.get('flags_equal'), p.get('within_tolerance'), p.get('max_err_over_tolerance'))
except Exception as e:
    print('$v FAILED', e); print(open('/tmp/ab.err').read()[-600:])" | tee -a /tmp/ab_all.txt
done
done
cp /tmp/orig.so $LIB
python - <<'PY'
import collections, statistics
d = collections.defaultdict(list)
for line in open('/tmp/ab_all.txt'):
    p = line.split()
    if len(p) > 2 and p[1] != 'FAILED':
        d[p[0]].append(float(p[1]))
print('---- summary (G env-steps/s): median, best, n')
for k, v in sorted(d.items()):
    print('%-28s %.3f  %.3f  %d' % (k, statistics.median(v), max(v), len(v)))
PY
