#!/bin/bash
# latency floor (N = 1024: less than one CTA per SM) and loaded step time (N = 16384) of every variants/*.so
LIB=atc_reinforcement_learning_b200/csrc/libatc_b200.so
cp $LIB /tmp/orig.so
for v in variants/*.so; do
  cp $v $LIB
  echo "== $v"
  python tools/scale_probe.py 2>&1 | grep -E "N +(1024|8192|16384) "
done
cp /tmp/orig.so $LIB
