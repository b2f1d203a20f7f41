#!/bin/bash
mkdir -p gpurun_out
{
SWEEP_SIZES=13,14,23,24 timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"
timeout 300 python tools/kprof_lde.py 2>&1 | tail -8
} > gpurun_out/ab_run32.log 2>&1
cat gpurun_out/ab_run32.log
