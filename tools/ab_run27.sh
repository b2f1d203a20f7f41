#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
TAG=r02l
tools/ncu_lde.sh $TAG > /dev/null 2>&1
H=24 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tip5_hash10_kernel -s 2 -c 2 -f -o gpurun_out/${TAG}_tip5 python tools/run_once.py merkle 1 > gpurun_out/${TAG}_ncu_tip5.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_tip5.ncu-rep > gpurun_out/${TAG}_ncu_tip5_hash10_merkle24_summary.txt 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_tip5.ncu-rep 2>/dev/null | head -5
cut -c1-220 gpurun_out/${TAG}_ncu_tip5_hash10_merkle24_summary.txt | head -20
cut -c1-300 gpurun_out/${TAG}_ncu_lde26_summary.txt | head -12
rm -f gpurun_out/*.ncu-rep
