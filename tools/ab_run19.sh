#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bfe_ntt_matches_oracle or xfe_ntt_matches_oracle or batched_ntt or four_pass or near_p or aligned_view" 2>&1 | tail -3
for m in 0 0x3e0; do echo "== TF21_MID_MASK=$m"; SWEEP_SIZES=15,16,17,18,19,25,26,27 TF21_MID_MASK=$m timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"; done
} > gpurun_out/ab_run19.log 2>&1
