#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lde or coset or poly or config3 or extrapolate" > gpurun_out/s4_tests8.log 2>&1; tail -2 gpurun_out/s4_tests8.log
python tools/kprof_lde.py > gpurun_out/s4_kprof_lde.txt 2>&1; cat gpurun_out/s4_kprof_lde.txt
python tools/quick_bench.py lde 2>&1 | grep lde
