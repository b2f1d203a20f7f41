#!/bin/bash
mkdir -p gpurun_out
{
for lib in libtf21.so libtf21_occ.so; do echo "== $lib"; TF21_LIB=$PWD/twenty-first_b200/$lib SWEEP_SIZES=6,7,13,14,19,21,22,23,24 timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"; TF21_LIB=$PWD/twenty-first_b200/$lib timeout 300 python tools/quick_bench.py lde 2>&1 | grep lde; TF21_LIB=$PWD/twenty-first_b200/$lib timeout 300 python tools/kprof_lde.py 2>&1 | tail -8; done
} > gpurun_out/ab_run24.log 2>&1
