#!/bin/bash
mkdir -p gpurun_out
{
for lib in libtf21.so libtf21_nf.so libtf21.so libtf21_nf.so; do echo "== $lib"; TF21_LIB=$PWD/twenty-first_b200/$lib timeout 300 python tools/quick_bench.py merkle tip5 2>&1 | grep -E "merkle 2\^24|hash_10"; done
TF21_LIB=$PWD/twenty-first_b200/libtf21_nf.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tip5 or merkle_tree" 2>&1 | tail -2
} > gpurun_out/ab_run26.log 2>&1
