#!/bin/bash
mkdir -p gpurun_out
export TF21_LIB=$PWD/twenty-first_b200/libtf21_nt.so
{
for nt in 0 1; do for th in 4 8 12; do
echo "== NT=$nt threads=$th"; TF21_STAGE_NT=$nt TF21_STAGE_THREADS=$th timeout 300 python tools/e2e_pageable.py 128 2>&1 | grep -E "ring|par_new.*ring"
done; done
for kb in 512 1024 2048; do
echo "== NT=1 threads=8 piece=${kb}KB"; TF21_STAGE_PIECE_KB=$kb TF21_STAGE_THREADS=8 timeout 300 python tools/e2e_pageable.py 128 2>&1 | grep -E "ring|par_new.*ring"
done
} > gpurun_out/ab_run13.log 2>&1
