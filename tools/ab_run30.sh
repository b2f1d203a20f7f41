#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02m_pytest_gpu.txt; cat gpurun_out/r02m_pytest_gpu.txt
timeout 900 python tools/size_sweep.py > gpurun_out/r02m_size_sweep.txt 2>&1; grep "w=1" gpurun_out/r02m_size_sweep.txt | sed -n 9,24p
python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -c 600 gpurun_out/r02m_bench.json
