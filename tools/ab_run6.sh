#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 bash tools/ab_env.sh "" "TF21_TMA_AHEAD=740" "TF21_TMA_AHEAD=370" "TF21_TMA_AHEAD=1480" "TF21_TMA_AHEAD=148" "TF21_TMA_AHEAD=2960" > gpurun_out/s4_ab6.log 2>&1
cat gpurun_out/s4_ab6.log
