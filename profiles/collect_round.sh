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
python bench.py --precision fp64 --no-cpu-baseline --no-secondary > $O/${R}_bench_fp64.json 2>> $O/${R}_bench.err
python bench.py --workload m12l24 --no-cpu-baseline --no-secondary > $O/${R}_bench_m12l24.json 2>> $O/${R}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
python profiles/probe_sweep.py --all > $O/${R}_probe_sweep.txt 2>&1
python profiles/probe_mix.py > $O/${R}_probe_mix.txt 2>&1
python profiles/latency.py > $O/${R}_latency.txt 2>&1
python profiles/small_batch.py > $O/${R}_small_batch.txt 2>&1
python profiles/phase_times.py > $O/${R}_phase_times.txt 2>&1
python profiles/collapse_batch.py > $O/${R}_collapse_batch.txt 2>&1
python profiles/potential_batch.py > $O/${R}_potential_batch.txt 2>&1
python profiles/trajectory_stream.py > $O/${R}_trajectory_stream.txt 2>&1
python profiles/chain_sweep.py --configs 0x0,64x1,128x1,256x1 > $O/${R}_chain_sweep.txt 2>&1
python profiles/cluster_chain.py > $O/${R}_cluster_chain.txt 2>&1
# profiler passes: profiles/collect_ncu.sh
tail -2 $O/${R}_pytest_gpu.txt
cat $O/${R}_bench.json | python profiles/bench_line.py
