#!/bin/bash
mkdir -p gpurun_out
{
TF21_MID2_MASK=0x2a0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(bfe_ntt_matches_oracle and (15 or 17 or 19)) or (xfe_ntt_matches_oracle and (15 or 19)) or near_p or batched_ntt" 2>&1 | tail -3
for m in 0x280 0x2a0; do echo "== TF21_MID2_MASK=$m"; SWEEP_SIZES=15,17,19,25 TF21_MID2_MASK=$m timeout 600 python tools/size_sweep.py 2>&1 | grep "w="; done
} > gpurun_out/ab_run31.log 2>&1
