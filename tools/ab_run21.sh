#!/bin/bash
mkdir -p gpurun_out
{
echo "== libtf21_c.so (6/5/4 CTAs) mask 0x3e0"; TF21_MID_MASK=0x3e0 TF21_LIB=$PWD/twenty-first_b200/libtf21_c.so SWEEP_SIZES=15,16,17,18,19,26 timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"
for pf in 0 400 800 1600 3200; do echo "== libtf21_b.so (5/4/3 CTAs) prefetch $pf"; TF21_MID_PREFETCH=$pf TF21_MID_MASK=0x3e0 TF21_LIB=$PWD/twenty-first_b200/libtf21_b.so SWEEP_SIZES=15,16,17,18,19,26 timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"; done
} > gpurun_out/ab_run21.log 2>&1
