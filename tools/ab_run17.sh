#!/bin/bash
mkdir -p gpurun_out
{
for m in 0 0x3e0; do echo "== TF21_MID_MASK=$m"; SWEEP_SIZES=15,16,17,18,19,25,26,27 TF21_MID_MASK=$m timeout 600 python tools/size_sweep.py 2>&1 | grep "w=1"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bfe_ntt_matches_oracle or xfe_ntt_matches_oracle or batched_ntt or four_pass" 2>&1 | tail -2
} > gpurun_out/ab_run17.log 2>&1
