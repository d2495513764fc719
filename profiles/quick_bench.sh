#!/bin/bash
# the four one-line device-resident numbers used while tuning (run through gpurun)
B="python bench.py --no-cpu-baseline --no-secondary"
echo "fp32  cube1000:"; $B | python profiles/bench_line.py
echo "fp64  cube1000:"; $B --precision fp64 | python profiles/bench_line.py
echo "fp32  m12l24:"; $B --workload m12l24 | python profiles/bench_line.py
echo "fp64  m12l24:"; $B --workload m12l24 --precision fp64 | python profiles/bench_line.py
