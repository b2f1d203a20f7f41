#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 bash tools/ab_libs.sh twenty-first_b200/libtf21.so twenty-first_b200/ab/v_mb4.so twenty-first_b200/ab/v_mb6.so > gpurun_out/s4_ab7.log 2>&1
cat gpurun_out/s4_ab7.log
