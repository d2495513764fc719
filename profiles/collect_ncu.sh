#!/bin/bash
# ncu evidence of a round (run through gpurun, ONE call):  gpurun --timeout 1500 -- 'bash profiles/collect_ncu.sh r02'
# Every profiled command first exits 0 without ncu (the `&&` chain); reports are summarised on the box
# and dropped (gpurun merges back at most 64 MiB), except the fp32 one which is kept for source-level reading.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
$B > $O/${R}_plain_fp32.json 2> $O/${R}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv $B > $O/${R}_ncu_launches.log 2>&1
$B > /dev/null 2>> $O/${R}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:mch_optimize -s 3 -c 1 -f -o $O/${R}_full_fp32 $B > $O/${R}_ncu_full_fp32.log 2>&1
# fp64: the full-evaluation kernel (early rejection, the mode's default, has its own captures: ncu_early.sh)
$B --precision fp64 --early-reject off > /dev/null 2>> $O/${R}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:mch_optimize -s 3 -c 1 -f -o $O/${R}_full_fp64 $B --precision fp64 --early-reject off > $O/${R}_ncu_full_fp64.log 2>&1
python profiles/collapse_batch.py 1184 > $O/${R}_plain_collapse.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mch_collapse -s 1 -c 1 -f -o $O/${R}_full_collapse_fp64 python profiles/collapse_batch.py 1184 > $O/${R}_ncu_collapse.log 2>&1
python profiles/collapse_batch.py 1184 fp32 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mch_collapse -s 1 -c 1 -f -o $O/${R}_full_collapse_fp32 python profiles/collapse_batch.py 1184 fp32 > $O/${R}_ncu_collapse32.log 2>&1
python profiles/potential_batch.py > $O/${R}_plain_potential.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mch_potential -s 1 -c 1 -f -o $O/${R}_full_potential python profiles/potential_batch.py > $O/${R}_ncu_potential.log 2>&1
for k in fp32 fp64 collapse_fp64 collapse_fp32 potential; do
  if [ -f $O/${R}_full_$k.ncu-rep ]; then
    python profiles/ncu_summary.py $O/${R}_full_$k.ncu-rep --json $O/${R}_ncu_full_$k.json > $O/${R}_ncu_full_$k.txt 2>&1
    ncu -i $O/${R}_full_$k.ncu-rep --page source --csv > $O/${R}_source_$k.csv 2>/dev/null
    gzip -f $O/${R}_source_$k.csv
    rm -f $O/${R}_full_$k.ncu-rep
  fi
done
ls -la $O | tail -30
