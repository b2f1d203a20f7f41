#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests9.log 2>&1; tail -3 gpurun_out/s4_tests9.log
python tools/size_sweep.py > gpurun_out/s4_size_sweep.txt 2>&1; cat gpurun_out/s4_size_sweep.txt
