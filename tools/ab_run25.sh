#!/bin/bash
mkdir -p gpurun_out
{
for lib in libtf21_base.so libtf21_dg.so libtf21_base.so libtf21_dg.so; do echo "== $lib"; TF21_LIB=$PWD/twenty-first_b200/$lib timeout 300 python tools/quick_bench.py merkle tip5 2>&1 | grep -E "merkle|hash_10"; done
TF21_LIB=$PWD/twenty-first_b200/libtf21_dg.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tip5 or merkle or mmr or authentication or sample" 2>&1 | tail -2
} > gpurun_out/ab_run25.log 2>&1
