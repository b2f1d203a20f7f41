#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests_opt.log 2>&1
tail -3 gpurun_out/s4_tests_opt.log
timeout 900 bash tools/ab_libs.sh twenty-first_b200/ab/base.so twenty-first_b200/libtf21.so > gpurun_out/s4_ab3.log 2>&1
cat gpurun_out/s4_ab3.log
python tools/quick_bench.py ntt lde > gpurun_out/s4_quick3.log 2>&1; cat gpurun_out/s4_quick3.log
