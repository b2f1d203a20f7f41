#!/bin/bash
mkdir -p gpurun_out
export TF21_LIB=$PWD/twenty-first_b200/libtf21_k9.so
{
TF21_MID_MASK=0x3e0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(bfe_ntt_matches_oracle and (19 or 18)) or (xfe_ntt_matches_oracle and 19) or near_p or aligned_view" 2>&1 | tail -3
for m in 0x1e0 0x3e0; do echo "== TF21_MID_MASK=$m"; SWEEP_SIZES=18,19 TF21_MID_MASK=$m timeout 600 python tools/size_sweep.py 2>&1 | grep "w="; done
} > gpurun_out/ab_run28.log 2>&1
