#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (memcheck, then racecheck on the shared-memory kernels)
cd /root/repo; mkdir -p gpurun_out
TAG=${1:-r02j}
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/${TAG}_memcheck.raw python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "near_p or tma_tile or small_n or (bfe_ntt and (11 or 15 or 16 or 17 or 18 or 19 or 20 or 21)) or xfe_ntt_matches_oracle and 10 or lde and not full or merkle_tree_matches or tip5_reference or mmr or authentication or aligned_view or (pinned_ring and 5000)" > gpurun_out/${TAG}_memcheck.out 2>&1
(echo "========= COMPUTE-SANITIZER memcheck"; tail -3 gpurun_out/${TAG}_memcheck.out; grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/${TAG}_memcheck.raw | head -20) > gpurun_out/${TAG}_compute_sanitizer_memcheck.log
cat gpurun_out/${TAG}_compute_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/${TAG}_racecheck.raw python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "near_p and (10 or 12 or 16) or (bfe_ntt and (16 or 17 or 19)) or tma_tile or small_n or merkle_tree_matches and (5 or 9 or 12)" > gpurun_out/${TAG}_racecheck.out 2>&1
(echo "========= COMPUTE-SANITIZER racecheck"; tail -3 gpurun_out/${TAG}_racecheck.out; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_racecheck.raw | head -20) > gpurun_out/${TAG}_compute_sanitizer_racecheck.log
cat gpurun_out/${TAG}_compute_sanitizer_racecheck.log
