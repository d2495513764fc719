set -u
O=gpurun_out; R=r02b
python profiles/collapse_batch.py 1184 fp64 > $O/${R}_plain_collapse.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mch_collapse -s 1 -c 1 -f -o $O/${R}_full_collapse_fp64 python profiles/collapse_batch.py 1184 fp64 > $O/${R}_ncu_collapse.log 2>&1
python profiles/collapse_batch.py 1184 fp32 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mch_collapse -s 1 -c 1 -f -o $O/${R}_full_collapse_fp32 python profiles/collapse_batch.py 1184 fp32 > $O/${R}_ncu_collapse32.log 2>&1
for k in collapse_fp64 collapse_fp32; do
  python profiles/ncu_summary.py $O/${R}_full_$k.ncu-rep --json $O/${R}_ncu_full_$k.json > $O/${R}_ncu_full_$k.txt 2>&1
  ncu -i $O/${R}_full_$k.ncu-rep --page source --csv > $O/${R}_source_$k.csv 2>/dev/null
  gzip -f $O/${R}_source_$k.csv; rm -f $O/${R}_full_$k.ncu-rep
done
cat $O/${R}_ncu_full_collapse_fp32.txt
