set -u
O=gpurun_out
echo "== probe sweep, mix on" ; python profiles/probe_sweep.py --all
echo "== probe sweep, mix off"; MCH_PROBES_LIB=variants/libprobes_nomix.so python profiles/probe_sweep.py --all
echo "== bench mix on"; python bench.py --no-cpu-baseline --no-secondary | python profiles/bench_line.py
echo "== bench mix off"; MCHAMMER_B200_LIB=variants/libmch_nomix.so python bench.py --no-cpu-baseline --no-secondary | python profiles/bench_line.py
echo "== m12l24 mix on"; python bench.py --no-cpu-baseline --no-secondary --workload m12l24 | python profiles/bench_line.py
echo "== m12l24 mix off"; MCHAMMER_B200_LIB=variants/libmch_nomix.so python bench.py --no-cpu-baseline --no-secondary --workload m12l24 | python profiles/bench_line.py
python -m pytest tests -m gpu -q -x --timeout 900 -k "fp32 or fast or tolerance or cluster or threads" 2>&1 | tail -5
