#!/bin/bash
# ncu --set full of the two kernels of the 256-column 2^20 batch (one launch each, after a warm-up pair)
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-s4}
COLS=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt1024 -s 2 -c 2 -f -o gpurun_out/${TAG}_ntt20_cols256 python tools/run_once.py ntt20 2 > gpurun_out/${TAG}_ncu_ntt.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_ntt.log
python tools/ncu_summary.py gpurun_out/${TAG}_ntt20_cols256.ncu-rep > gpurun_out/${TAG}_ncu_ntt1024_cols256_summary.txt 2>&1
cat gpurun_out/${TAG}_ncu_ntt1024_cols256_summary.txt
