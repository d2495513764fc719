#!/bin/bash
# Tuning build of the step kernel only:  bash profiles/quick_variant.sh NAME [f32|f64|both] [-DFOO=1 ...]
# compiles csrc/mch_optimize_f32.cu / _f64.cu with -DMCH_QUICK (mu = 3, f64 position arrays) plus the given
# defines and links them with the objects of the regular build -> variants/libmch_NAME.so
# (use with MCHAMMER_B200_LIB=variants/libmch_NAME.so, profiles/ab_variants.sh).
set -e
NAME=$1; WHICH=${2:-both}; shift; shift || true
D=mchammer_b200; B=$D/build; V=variants/obj_$NAME
mkdir -p $V variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -fmad=false -Xcompiler -fPIC -DMCH_QUICK $*"
pids=()
for k in f32 f64; do
  if [ "$WHICH" = both ] || [ "$WHICH" = $k ]; then
    nvcc $FLAGS -Xptxas -v -c -o $V/mch_optimize_$k.o $D/csrc/mch_optimize_$k.cu > $V/ptxas_$k.txt 2>&1 &
    pids+=($!)
  else
    cp $B/mch_optimize_$k.o $V/mch_optimize_$k.o
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libmch_$NAME.so $B/mch_api.o $V/mch_optimize_f32.o $V/mch_optimize_f64.o $B/mch_wide_f32.o $B/mch_wide_f64.o
grep -h "spill\|error" $V/ptxas_*.txt | sort | uniq -c | sort -rn | head -8
ls -la variants/libmch_$NAME.so
