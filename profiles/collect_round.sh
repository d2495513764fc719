#!/bin/bash
# Re-collect the single-GPU evidence of a round on the GPU box (run through gpurun):
#   gpurun --timeout 1500 -- 'bash profiles/collect_round.sh r01'
# Everything is written to gpurun_out/<round>_*; copy what should be judged into profiles/.
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.txt 2>&1
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err
python bench.py --precision fp64 --no-cpu-baseline > $O/${R}_bench_fp64.json 2>> $O/${R}_bench.err
python bench.py --workload m12l24 --no-cpu-baseline > $O/${R}_bench_m12l24.json 2>> $O/${R}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
python profiles/probe_sweep.py --all > $O/${R}_probe_sweep.txt 2>&1
python profiles/probe_mix.py > $O/${R}_probe_mix.txt 2>&1
python profiles/latency.py > $O/${R}_latency.txt 2>&1
python profiles/small_batch.py > $O/${R}_small_batch.txt 2>&1
python profiles/phase_times.py > $O/${R}_phase_times.txt 2>&1
python profiles/collapse_batch.py > $O/${R}_collapse_batch.txt 2>&1
python profiles/potential_batch.py > $O/${R}_potential_batch.txt 2>&1
python profiles/trajectory_stream.py > $O/${R}_trajectory_stream.txt 2>&1
# profiler passes: the same command exited 0 above without ncu
if python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_bench_short.json 2>> $O/${R}_bench.err; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:mch_optimize -c 1 -f -o $O/${R}_full_fp32 \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_ncu_full_fp32.log 2>&1
fi
if python bench.py --steps 2 --warmup 1 --no-cpu-baseline --precision fp64 > /dev/null 2>> $O/${R}_bench.err; then
  ncu --set full --clock-control none --import-source on -k regex:mch_optimize -c 1 -f -o $O/${R}_full_fp64 \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --precision fp64 > $O/${R}_ncu_full_fp64.log 2>&1
fi
# summarise the full captures here and drop the reports (gpurun merges back at most 64 MiB)
for k in fp32 fp64; do
  if [ -f $O/${R}_full_$k.ncu-rep ]; then
    python profiles/ncu_summary.py $O/${R}_full_$k.ncu-rep --json $O/${R}_ncu_full_$k.json > $O/${R}_ncu_full_$k.txt 2>&1
    rm -f $O/${R}_full_$k.ncu-rep
  fi
done
tail -2 $O/${R}_pytest_gpu.txt
cat $O/${R}_bench.json | python profiles/bench_line.py
