#!/bin/bash
mkdir -p gpurun_out
{
tools/ab_env.sh "TF21_L2_GROUP=0" "TF21_L2_GROUP=2" "TF21_L2_GROUP=4" "TF21_L2_GROUP=6" "TF21_L2_GROUP=8" "TF21_L2_GROUP=16" "TF21_L2_GROUP=32" "TF21_L2_GROUP=0"
TF21_L2_GROUP=4 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "config1_full or ntt_matches_oracle or round_trip" 2>&1 | tail -3
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"
timeout 600 python tools/e2e_pageable.py 256
} > gpurun_out/ab_run11.log 2>&1
