#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bfe_ntt_matches_oracle or xfe_ntt_matches_oracle or batched_ntt or four_pass or near_p or aligned_view or poly_ or coset" 2>&1 | tail -4
for m in 0 0x3e0; do echo "== TF21_MID_MASK=$m"; SWEEP_SIZES=14,15,16,17,18,19,20,24,25,26 TF21_MID_MASK=$m timeout 600 python tools/size_sweep.py; done
} > gpurun_out/ab_run16.log 2>&1
