#!/bin/bash
mkdir -p gpurun_out
{
for lib in libtf21.so libtf21_b.so; do echo "== $lib mask 0x3e0"; TF21_MID_MASK=0x3e0 TF21_LIB=$PWD/twenty-first_b200/$lib SWEEP_SIZES=15,16,17,18,19,26 timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bfe_ntt_matches_oracle or xfe_ntt_matches_oracle or batched_ntt or four_pass or near_p or aligned_view" 2>&1 | tail -2
} > gpurun_out/ab_run20.log 2>&1
