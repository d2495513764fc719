#!/bin/bash
# ncu capture of the early-rejection instantiation of the step kernel (fp32, bench workload), through gpurun:
#   gpurun --timeout 900 -- 'bash profiles/ncu_early.sh r02'
set -u
R=${1:-r02}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --early-reject on"
$B > $O/${R}_plain_early_fp32.json 2> $O/${R}_plain_early.err &&
ncu --set full --clock-control none --import-source on -k regex:mch_optimize -s 3 -c 1 -f -o $O/${R}_full_early_fp32 $B > $O/${R}_ncu_full_early_fp32.log 2>&1
if [ -f $O/${R}_full_early_fp32.ncu-rep ]; then
  python profiles/ncu_summary.py $O/${R}_full_early_fp32.ncu-rep --json $O/${R}_ncu_full_early_fp32.json > $O/${R}_ncu_full_early_fp32.txt 2>&1
  ncu -i $O/${R}_full_early_fp32.ncu-rep --page source --csv > $O/${R}_source_early_fp32.csv 2>/dev/null
  gzip -f $O/${R}_source_early_fp32.csv
  rm -f $O/${R}_full_early_fp32.ncu-rep
fi
ls -la $O | tail -8
