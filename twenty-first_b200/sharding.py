"""Host-side sharding logic for the multi-GPU path (one process per GPU).

The hot path shards with no data-path collective:
  * NTT columns are independent (the reference never splits one NTT across threads either,
    ntt.rs:250-269): rank g of G owns columns [g*cols/G, (g+1)*cols/G).
  * Merkle: rank g builds the subtree over leaves [g*n/G, (g+1)*n/G) -- a contiguous power-of-two
    block of every level of the heap-indexed tree (cf. `subtrees_mut`, merkle_tree.rs:247-275) --
    then ONE all-gather of the G local roots (40 bytes per rank, the tree cap) and every rank
    finishes the top log2(G) levels.

Everything here is index algebra plus the collective; the hashing itself is injected
(`build_tree`), so the same code runs over `device.merkle_build` (CUDA, NCCL) in bench.py and over
a CPU checker with gloo in tests/test_sharding_gloo.py.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def is_pow2(x: int) -> bool:
    return x > 0 and (x & (x - 1)) == 0


def column_shard(n_columns: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the columns owned by `rank`; remainder columns go to the first ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_columns, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def leaf_shard(n_leafs: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the leaves whose subtree `rank` builds. Requires power-of-two n_leafs and
    world (a subtree must be a complete binary tree) and at least one leaf per rank."""
    if not (is_pow2(n_leafs) and is_pow2(world)):
        raise ValueError("n_leafs and world size must be powers of two")
    if n_leafs < world:
        raise ValueError("fewer leaves than ranks")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per = n_leafs // world
    return rank * per, (rank + 1) * per


def global_node_index(local_index: int, n_local_leafs: int, shard: int, n_shards: int) -> int:
    """Heap index in the global tree of node `local_index` (>= 1) of shard `shard`'s local tree.
    Same mapping as merkle_scatter_kernel (csrc/tip5_kernels.cuh)."""
    if local_index < 1 or local_index >= 2 * n_local_leafs:
        raise ValueError("local node index out of range")
    level = local_index.bit_length() - 1          # local node in [2^level, 2^(level+1))
    width = 1 << level
    return n_shards * width + shard * width + (local_index - width)


def scatter_subtree(local_nodes: np.ndarray, shard: int, n_shards: int, global_nodes: np.ndarray) -> None:
    """numpy version of tf21_merkle_scatter_subtree_dev (host reference for the tests)."""
    n_local = local_nodes.shape[0] // 2
    for level in range(n_local.bit_length()):
        width = 1 << level
        if width > n_local:
            break
        g0 = n_shards * width + shard * width
        global_nodes[g0:g0 + width] = local_nodes[width:2 * width]


def sharded_merkle_root(local_leafs, rank: int, world: int, build_tree: Callable, all_gather_roots: Callable):
    """Root of the tree over the concatenation of all ranks' leaves.

    build_tree(leafs) -> heap-indexed node array (2 * len(leafs) digests, [1] = root)
    all_gather_roots(root) -> array of the `world` roots in rank order (the only collective)
    Returns (root, local_nodes, cap_nodes); cap_nodes is the heap-indexed tree over the roots.
    """
    local_nodes = build_tree(local_leafs)
    roots = all_gather_roots(local_nodes[1])
    if world == 1:
        return local_nodes[1], local_nodes, None
    cap_nodes = build_tree(roots)
    return cap_nodes[1], local_nodes, cap_nodes
