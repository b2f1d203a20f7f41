"""Developer timing helper: tf21_ntt / tf21_intt on HOST slices -- pinned memory against plain (pageable) numpy memory,
the latter through the pinned staging ring (csrc/host_stage.cuh) and, with TF21_NO_STAGE_RING=1, through the driver's
own pageable copies.  Usage: python tools/e2e_pageable.py [cols]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
api = importlib.import_module("twenty-first_b200.api")
tf.device.init(0)
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = 1 << 20
rng = np.random.default_rng(1)
src = rng.integers(0, 0xFFFFFFFF00000001, size=cols * n, dtype=np.uint64)


def run(buf, label):
    buf[:] = src
    api.ntt_batch(buf, n, 1, False)  # warm-up (tables, ring)
    api.ntt_batch(buf, n, 1, True)
    assert np.array_equal(buf, src), label + ": round trip differs"
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        api.ntt_batch(buf, n, 1, False)
        api.ntt_batch(buf, n, 1, True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    assert np.array_equal(buf, src), label + ": round trip differs"
    gbs = 2 * cols * n * 8 / best / 1e9
    print(f"{label}: {2 * cols / best:.0f} NTT/s  ({best * 1e3:.1f} ms per forward+inverse of {cols} columns, "
          f"{gbs:.1f} GB/s in each direction)", flush=True)
    return buf


pinned = torch.empty(cols * n, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
run(pinned, "pinned")
fwd_ref = pinned.copy()
api.ntt_batch(fwd_ref, n, 1, False)
for th in ([os.environ["TF21_STAGE_THREADS"]] if "TF21_STAGE_THREADS" in os.environ else ["2", "4", "6", "8"]):
    os.environ["TF21_STAGE_THREADS"] = th
    pageable = np.empty(cols * n, dtype=np.uint64)
    run(pageable, f"pageable via ring, {th}+{th} threads")
    api.ntt_batch(pageable, n, 1, False)
    assert np.array_equal(pageable, fwd_ref), "staged forward transform differs from the pinned path"
os.environ["TF21_NO_STAGE_RING"] = "1"
run(np.empty(cols * n, dtype=np.uint64), "pageable via driver copies")

# ---- host Merkle build (tf21_merkle_build on pageable slices) ----
del os.environ["TF21_NO_STAGE_RING"]
nl = 1 << 24
leafs = rng.integers(0, 0xFFFFFFFF00000001, size=5 * nl, dtype=np.uint64).reshape(nl, 5)
for label, env in (("ring", None), ("driver copies", "1")):
    if env:
        os.environ["TF21_NO_STAGE_RING"] = env
    tf.MerkleTree.par_new(leafs)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        tree = tf.MerkleTree.par_new(leafs)
        best = min(best, time.perf_counter() - t0)
    print(f"MerkleTree.par_new 2^24 host leaves via {label}: {best * 1e3:.1f} ms ({nl / best / 1e6:.1f} M leaves/s)", flush=True)
