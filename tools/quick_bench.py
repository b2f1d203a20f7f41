"""Developer timing helper (not the contract bench): device-resident kernels timed with CUDA events."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
dev = tf.device
dev.init(0)
cuda = torch.device("cuda:0")
HBM = 6448.1e9


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def rand_words(n):
    x = torch.randint(0, 2**62, (n,), dtype=torch.int64, device=cuda)
    return x


what = sys.argv[1:] or ["ntt", "merkle", "tip5", "rows", "lde"]
if "ntt" in what:
    for log2n, batch in ((10, 65536), (16, 1024), (20, 64), (20, 256), (24, 8)):
        x = rand_words((1 << log2n) * batch)
        best, med = timeit(lambda: dev.ntt_(x, 1 << log2n, 1, False))
        bytes_ = 16 * (1 << log2n) * batch
        print(f"ntt 2^{log2n} x{batch}: best {best:.3f} ms  med {med:.3f} ms  {batch/best*1e3:.0f} NTT/s  "
              f"{bytes_/best/1e6:.0f} GB/s = {bytes_/best*1e3/HBM*100:.1f}% of HBM roofline")
        del x
if "merkle" in what:
    for h in (16, 20, 24):
        n = 1 << h
        leafs = rand_words(5 * n)
        nodes = torch.zeros(10 * n, dtype=torch.int64, device=cuda)
        best, med = timeit(lambda: dev.merkle_build(leafs, nodes))
        print(f"merkle 2^{h}: best {best:.3f} ms  {n/best*1e3/1e9:.3f} G leaves/s  "
              f"{80*n/best*1e3/HBM*100:.2f}% of HBM roofline")
        del leafs, nodes
if "tip5" in what:
    n = 1 << 22
    inp = rand_words(10 * n)
    out = torch.zeros(5 * n, dtype=torch.int64, device=cuda)
    best, med = timeit(lambda: dev.tip5_hash_10(inp, out))
    print(f"hash_10 x2^22: best {best:.3f} ms  {n/best*1e3/1e9:.3f} G hash/s")
if "rows" in what:
    n_rows, n_cols = 1 << 20, 100
    cols = rand_words(n_rows * n_cols)
    out = torch.zeros(5 * n_rows, dtype=torch.int64, device=cuda)
    best, med = timeit(lambda: dev.tip5_hash_columns(cols, n_rows, n_cols, out), iters=3, warm=1)
    perms = n_rows * ((n_cols + 1 + 9) // 10)
    print(f"hash_columns 2^20 rows x {n_cols} cols: best {best:.3f} ms  {perms/best*1e3/1e9:.3f} G permutations/s  "
          f"{8*n_rows*n_cols/best/1e6:.0f} GB/s read")
if "lde" in what:
    for li, lo in ((18, 22), (22, 26)):
        vals = rand_words(3 << li)
        out = torch.zeros(3 << lo, dtype=torch.int64, device=cuda)
        g = tf.BFieldElement.generator()
        best, med = timeit(lambda: dev.coset_lde(vals, 3, g, 1 << lo, g, out), iters=3, warm=1)
        bytes_ = 24 * ((1 << li) + (1 << lo))
        print(f"lde xfe 2^{li}->2^{lo}: best {best:.3f} ms  {bytes_/best*1e3/HBM*100:.1f}% of HBM roofline")
print("launches", dev.kernel_launch_count())
