#!/bin/bash
# A/B timing of kernel variants: every variants/*.so is copied over the in-tree library and benched (device-resident
# timing only).  Usage (on the GPU box): bash tools/ab_variants.sh [bench args]
LIB=atc_reinforcement_learning_b200/csrc/libatc_b200.so
cp $LIB /tmp/orig.so
for rep in 1 2; do
for v in variants/*.so; do
  cp $v $LIB
  python bench.py --steps 16384 --warmup 1024 --skip-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,3), round(d['roofline']['frac'],4))"
done
done
cp /tmp/orig.so $LIB
