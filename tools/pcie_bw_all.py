"""Host<->device ceiling of one box and the single-process sharded NTT against it (developer helper / evidence).

  python tools/pcie_bw_all.py            # every visible GPU

1. pinned-memory H2D, D2H and simultaneous H2D+D2H bandwidth of each GPU alone and of 1/2/4/8 GPUs at once
   (one stream pair per GPU, all driven from this process) -- the ceiling any host-slice entry point can reach;
2. tf21_ntt_sharded (host slices, one worker thread per shard, include/tf21.h) on 1/2/4/8 shards with the same
   256 x 2^20 columns per shard as bench.py's e2e leg;
3. the topology the driver reports (nvidia-smi topo -m) so that NUMA placement can be read next to the numbers.
"""
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

n_gpu = torch.cuda.device_count()
N = 1 << 30  # 1 GiB per direction per GPU
res = {"gpus": n_gpu}


def bw(devs, mode, reps=3):
    bufs = []
    for d in devs:
        torch.cuda.set_device(d)
        h1 = torch.empty(N, dtype=torch.uint8).pin_memory()
        h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
        d1 = torch.empty(N, dtype=torch.uint8, device=f"cuda:{d}")
        d2 = torch.empty(N, dtype=torch.uint8, device=f"cuda:{d}")
        bufs.append((d, h1, h2, d1, d2, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))

    def go():
        for d, h1, h2, d1, d2, s1, s2 in bufs:
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d1.copy_(h1, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)

    def sync():
        for d in devs:
            torch.cuda.synchronize(d)

    go()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    sync()
    dt = (time.perf_counter() - t0) / reps
    return len(devs) * N / dt / 1e9  # GB/s per direction, summed over the GPUs


counts = [c for c in (1, 2, 4, 8) if c <= n_gpu]
res["pinned_copy_GBps_per_direction_aggregate"] = {
    str(c): {m: round(bw(list(range(c)), m), 1) for m in ("h2d", "d2h", "both")} for c in counts}
if n_gpu > 1:
    res["each_gpu_alone_both_GBps"] = [round(bw([d], "both"), 1) for d in range(n_gpu)]

tf = importlib.import_module("twenty-first_b200")
tf.check(tf.lib.tf21_init(0))
n, cols = 1 << 20, 256
sh = {}
for c in counts:
    host = torch.empty(c * cols * n, dtype=torch.int64).pin_memory()
    host.random_(0, 2**62)
    p = host.data_ptr()
    tf.check(tf.lib.tf21_ntt_sharded(p, n, 1, c * cols, 0, c))  # warm-up: tables, pools
    tf.check(tf.lib.tf21_ntt_sharded(p, n, 1, c * cols, 1, c))
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        tf.check(tf.lib.tf21_ntt_sharded(p, n, 1, c * cols, 0, c))
        tf.check(tf.lib.tf21_ntt_sharded(p, n, 1, c * cols, 1, c))
    dt = (time.perf_counter() - t0) / reps
    sh[str(c)] = {"ntt_per_s": round(2 * c * cols / dt, 1), "GBps_each_direction": round(2 * c * cols * n * 8 / dt / 1e9, 1)}
    del host
res["tf21_ntt_sharded_host_slices"] = sh
try:
    res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
except Exception as e:  # noqa: BLE001
    res["topo"] = str(e)
res["host_cpus"] = len(os.sched_getaffinity(0))
print(json.dumps(res, indent=1))
