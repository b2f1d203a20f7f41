#!/bin/bash
mkdir -p gpurun_out
{
TF21_MID7_TWO_THREAD=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(bfe_ntt_matches_oracle and (17 or 19)) or (xfe_ntt_matches_oracle and (17 or 19)) or near_p or batched_ntt or four_pass" 2>&1 | tail -3
for t in 0 1; do echo "== TF21_MID7_TWO_THREAD=$t"; SWEEP_SIZES=17,19,27 TF21_MID7_TWO_THREAD=$t timeout 600 python tools/size_sweep.py 2>&1 | grep "w="; done
} > gpurun_out/ab_run29.log 2>&1
