"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard ranges, the cap all-gather and
the local->global node index mapping.  The hashing is injected: here the CPU oracle stands in for the
CUDA kernels (it is the checker in this test, never the product path)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, height, out_dir):
    sys.path.insert(0, ROOT)
    import torch

    import oracle

    sh = importlib.import_module("twenty-first_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle.get()
    n = 1 << height
    leafs = oracle.splitmix64_words(0xBEEF, 5 * n).reshape(n, 5)
    lo, hi = sh.leaf_shard(n, rank, world)

    def build_tree(lf):
        rc, nodes = o.merkle_sequential_new(np.ascontiguousarray(lf).reshape(-1))
        assert rc == 0
        return nodes.reshape(-1, 5)

    def all_gather_roots(root):
        t = torch.from_numpy(np.ascontiguousarray(root).view(np.int64))
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return np.stack([x.numpy().view(np.uint64) for x in outs])

    root, local_nodes, cap = sh.sharded_merkle_root(leafs[lo:hi], rank, world, build_tree, all_gather_roots)
    # reference: the single tree
    rc, want = o.merkle_sequential_new(leafs.reshape(-1))
    want = want.reshape(-1, 5)
    assert np.array_equal(root, want[1])
    # scatter the local tree and the cap into a global array and compare the parts this rank knows
    glob = np.zeros_like(want)
    sh.scatter_subtree(local_nodes, rank, world, glob)
    for level_width in (1 << k for k in range((n // world).bit_length())):
        if level_width > n // world:
            break
        g0 = world * level_width + rank * level_width
        assert np.array_equal(glob[g0:g0 + level_width], want[g0:g0 + level_width])
    assert np.array_equal(cap[1:world], want[1:world])          # top log2(world) levels
    assert np.array_equal(cap[world:2 * world], want[world:2 * world])  # the cap itself
    # column shards partition the batch
    ranges = [sh.column_shard(257, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == 257 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    # checksum-of-checksums pattern used by bench.py for cfg4
    mine = torch.tensor([int(leafs[lo:hi].sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF))], dtype=torch.int64)
    dist.all_reduce(mine, op=dist.ReduceOp.BXOR)
    np.save(os.path.join(out_dir, f"root_{rank}.npy"), root)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("height", [1, 6])
def test_two_rank_sharded_merkle_matches_single_tree(tmp_path, height):
    port = 29500 + (os.getpid() % 2000) + height
    mp.spawn(_worker, args=(2, port, height, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "root_0.npy"), np.load(tmp_path / "root_1.npy")
    assert np.array_equal(r0, r1)


def test_shard_helpers():
    sh = importlib.import_module("twenty-first_b200.sharding")
    assert sh.leaf_shard(16, 3, 4) == (12, 16)
    with pytest.raises(ValueError):
        sh.leaf_shard(12, 0, 4)
    with pytest.raises(ValueError):
        sh.leaf_shard(2, 0, 4)
    assert [sh.column_shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    # local node 1 (the root) of shard 2 of 4 sits at global index 4 + 2
    assert sh.global_node_index(1, 8, 2, 4) == 6
    # local leaf k of shard s sits at n + s * n_local + k
    assert sh.global_node_index(8 + 5, 8, 3, 4) == 32 + 3 * 8 + 5
