#!/bin/bash
# final single-GPU evidence of a round: full GPU suite, both bench arms + launch list, sanitizer, size sweep, ncu --set full of
# the 2^20 kernels, the mid pass and the Merkle top
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02l}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_pytest_gpu.txt
tools/evidence.sh $TAG
timeout 900 python tools/size_sweep.py > gpurun_out/${TAG}_size_sweep.txt 2>&1; tail -3 gpurun_out/${TAG}_size_sweep.txt
tools/ncu_ntt.sh $TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_mid_col -s 1 -c 1 -f -o gpurun_out/${TAG}_mid16 python tools/run_once.py ntt16 2 > gpurun_out/${TAG}_ncu_mid.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_mid16.ncu-rep > gpurun_out/${TAG}_ncu_mid_col_2_16_summary.txt 2>&1
H=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:merkle_top -s 1 -c 1 -f -o gpurun_out/${TAG}_mtop python tools/run_once.py merkle 2 > gpurun_out/${TAG}_ncu_mtop.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_mtop.ncu-rep > gpurun_out/${TAG}_ncu_merkle_top_2_16_summary.txt 2>&1
head -12 gpurun_out/${TAG}_ncu_mid_col_2_16_summary.txt | cut -c1-200; head -8 gpurun_out/${TAG}_ncu_merkle_top_2_16_summary.txt | cut -c1-200
tools/sanitize.sh $TAG
rm -f gpurun_out/*.ncu-rep
