#!/bin/bash
# bench.py on N GPUs of one box (run under `gpurun --gpus N`): tools/evidence_multi.sh N TAG
cd /root/repo; mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02j}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 2500 gpurun_out/${TAG}_bench_${N}gpu.json
if [ "$N" -ge 2 ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded" 2>&1 | tail -2; fi
