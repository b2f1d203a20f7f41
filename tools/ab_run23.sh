#!/bin/bash
mkdir -p gpurun_out
{
for lib in libtf21.so libtf21_mb5.so; do echo "== $lib"; TF21_LIB=$PWD/twenty-first_b200/$lib SWEEP_SIZES=3,5,6,7,8,9,10,11,12,21,22 timeout 600 python tools/size_sweep.py 2>&1 | grep "w="; TF21_LIB=$PWD/twenty-first_b200/$lib timeout 300 python tools/quick_bench.py lde 2>&1 | grep lde; done
} > gpurun_out/ab_run23.log 2>&1
