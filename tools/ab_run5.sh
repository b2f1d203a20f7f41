#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or lde or coset or poly or config" > gpurun_out/s4_tests5.log 2>&1; tail -2 gpurun_out/s4_tests5.log
timeout 1200 bash tools/ab_libs.sh twenty-first_b200/ab/opt.so twenty-first_b200/libtf21.so > gpurun_out/s4_ab5.log 2>&1
cat gpurun_out/s4_ab5.log
