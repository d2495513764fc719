#!/bin/bash
# Rebuild only csrc/mch_api.cu (potential + collapse kernels, C ABI) and relink the regular library.
set -e
D=mchammer_b200; B=$D/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -Xcompiler -fPIC -Xptxas -v "$@" -c -o $B/mch_api.o $D/csrc/mch_api.cu > $B/ptxas_api.txt 2>&1 || { tail -30 $B/ptxas_api.txt; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $D/libmchammer_b200.so $B/mch_api.o $B/mch_optimize_f32.o $B/mch_optimize_f64.o $B/mch_wide_f32.o $B/mch_wide_f64.o
grep -A1 "mch_collapse" $B/ptxas_api.txt | grep -v "^--" | grep -i "Used\|spill" | sort | uniq -c
