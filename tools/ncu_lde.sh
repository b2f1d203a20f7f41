#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02j}
timeout 900 ncu --set full --clock-control none --import-source on -s 6 -c 6 -f -o gpurun_out/${TAG}_lde26 python tools/run_once.py lde26 2 > gpurun_out/${TAG}_ncu_lde.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_lde.log
python tools/ncu_summary.py gpurun_out/${TAG}_lde26.ncu-rep > gpurun_out/${TAG}_ncu_lde26_summary.txt 2>&1
cat gpurun_out/${TAG}_ncu_lde26_summary.txt | cut -c1-400
