#!/bin/bash
# A/B of tuning variants built into variants/*.so:  bash profiles/ab_variants.sh name1 name2 ...
B="python bench.py --no-cpu-baseline --no-secondary"
for v in "$@"; do
  export MCHAMMER_B200_LIB=variants/libmch_$v.so
  echo "== $v"
  echo -n "fp32 cube1000 "; $B | python profiles/bench_line.py
  echo -n "fp64 cube1000 "; $B --precision fp64 | python profiles/bench_line.py
  echo -n "fp32 m12l24   "; $B --workload m12l24 | python profiles/bench_line.py
done
