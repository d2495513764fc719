B="python bench.py --no-cpu-baseline --no-secondary --precision fp64"
for v in "$@"; do
  export MCHAMMER_B200_LIB=variants/libmch_$v.so
  echo -n "== $v fp64 cube1000 "; $B | python profiles/bench_line.py
done
