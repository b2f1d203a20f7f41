#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "pinned_ring or leaf_half or config1_full or config2_merkle or merkle_tree_matches or coset_lde_full or poly_ or concurrent or sharded or shutdown" 2>&1 | tail -5
TF21_STAGE_THREADS=8 timeout 600 python tools/e2e_pageable.py 256
} > gpurun_out/ab_run12.log 2>&1
