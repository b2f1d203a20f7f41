#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ntt or coset or poly" 2>&1 | tail -2
for m in 0x1e0 0x1f8; do echo "== TF21_MID_MASK=$m"; SWEEP_SIZES=13,14,15,16,17,18,23,24,25,26 TF21_MID_MASK=$m timeout 600 python tools/size_sweep.py 2>&1 | grep "w="; done
TF21_MID_MASK=0x1f8 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bfe_ntt_matches_oracle or xfe_ntt_matches_oracle or batched_ntt or four_pass or near_p" 2>&1 | tail -2
} > gpurun_out/ab_run22.log 2>&1
