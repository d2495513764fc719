#!/bin/bash
# ncu captures of the early-rejection instantiation of the step kernel (bench workload, both precisions), through gpurun:
#   gpurun --timeout 900 -- 'bash profiles/ncu_early.sh r02'
set -u
R=${1:-r02}; O=gpurun_out; mkdir -p $O
for P in fp32 fp64; do
  B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --early-reject on --precision $P"
  $B > $O/${R}_plain_early_$P.json 2> $O/${R}_plain_early.err &&
  ncu --set full --clock-control none --import-source on -k regex:mch_optimize -s 3 -c 1 -f -o $O/${R}_full_early_$P $B > $O/${R}_ncu_full_early_$P.log 2>&1
  if [ -f $O/${R}_full_early_$P.ncu-rep ]; then
    python profiles/ncu_summary.py $O/${R}_full_early_$P.ncu-rep --json $O/${R}_ncu_full_early_$P.json > $O/${R}_ncu_full_early_$P.txt 2>&1
    ncu -i $O/${R}_full_early_$P.ncu-rep --page source --csv > $O/${R}_source_early_$P.csv 2>/dev/null
    gzip -f $O/${R}_source_early_$P.csv
    rm -f $O/${R}_full_early_$P.ncu-rep
  fi
done
ls -la $O | tail -8
