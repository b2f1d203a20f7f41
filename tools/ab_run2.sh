#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
TF21_COL_ORDER=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt and not small" > gpurun_out/s4_order_tests.log 2>&1
tail -3 gpurun_out/s4_order_tests.log
timeout 900 bash tools/ab_env.sh "" "TF21_COL_ORDER=0" "TF21_COL_ORDER=1" "TF21_COL_ORDER=2" "TF21_COL_ORDER=3" "TF21_COL_ORDER=5" > gpurun_out/s4_ab2.log 2>&1
cat gpurun_out/s4_ab2.log
