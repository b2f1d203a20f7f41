#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 1200 bash tools/ab_libs.sh twenty-first_b200/ab/opt.so $(ls twenty-first_b200/ab/v_*.so) > gpurun_out/s4_ab4.log 2>&1
cat gpurun_out/s4_ab4.log
