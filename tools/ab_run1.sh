#!/bin/bash
# one GPU visit: parity of the fused Tip5 build, then A/B timings
cd /root/repo
mkdir -p gpurun_out
TF21_LIB=$PWD/twenty-first_b200/ab/t5_fence.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tip5 or merkle or mmr or sample" > gpurun_out/s4_t5_tests.log 2>&1
tail -3 gpurun_out/s4_t5_tests.log
AB_MERKLE=1 timeout 900 bash tools/ab_libs.sh twenty-first_b200/ab/base.so twenty-first_b200/ab/t5_fence.so twenty-first_b200/ab/t5_nofence.so twenty-first_b200/ab/ntt_exitmid.so > gpurun_out/s4_ab1.log 2>&1
cat gpurun_out/s4_ab1.log
