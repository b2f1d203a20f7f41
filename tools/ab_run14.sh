#!/bin/bash
mkdir -p gpurun_out
{
for top in 0 2048 8192 16384 65536; do
echo "== TF21_MERKLE_TOP_CNT=$top"; TF21_MERKLE_TOP_CNT=$top timeout 300 python tools/quick_bench.py merkle 2>&1 | grep merkle
done
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
} > gpurun_out/ab_run14.log 2>&1
