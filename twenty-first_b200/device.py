"""Device-resident entry points: the measured path.  Buffers are torch CUDA tensors of dtype int64
(bit patterns of the raw u64 words); torch is only used for device memory and streams."""
from __future__ import annotations

import torch

from . import _binding as B


def _p(t: torch.Tensor):
    assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()
    return t.data_ptr() if t.numel() else None


def _stream():
    return torch.cuda.current_stream().cuda_stream


def init(device: int) -> None:
    """one process per GPU: select `device` for this thread and prepare it (tf21_set_device)"""
    B.check(B.lib.tf21_set_device(device))


def ntt_(data: torch.Tensor, n: int, width: int = 1, inverse: bool = False) -> None:
    batch = data.numel() // (n * width) if n else 0
    B.check(B.lib.tf21_ntt_dev(_p(data), n, width, batch, int(inverse), _stream()))


def coset_evaluate(coeffs: torch.Tensor, width: int, offset_raw: int, order: int, out: torch.Tensor) -> None:
    B.check(B.lib.tf21_coset_evaluate_dev(_p(coeffs), coeffs.numel() // width, width, offset_raw, order, _p(out),
                                          _stream()))


def coset_interpolate(values: torch.Tensor, width: int, offset_raw: int, out: torch.Tensor) -> None:
    B.check(B.lib.tf21_coset_interpolate_dev(_p(values), values.numel() // width, width, offset_raw, _p(out),
                                             _stream()))


def coset_lde(values: torch.Tensor, width: int, offset_in_raw: int, n_out: int, offset_out_raw: int,
              out: torch.Tensor) -> None:
    B.check(B.lib.tf21_coset_lde_dev(_p(values), values.numel() // width, offset_in_raw, n_out, offset_out_raw,
                                     width, _p(out), _stream()))


def tip5_permute_(states: torch.Tensor) -> None:
    B.check(B.lib.tf21_tip5_permute_dev(_p(states), states.numel() // 16, _stream()))


def tip5_hash_10(inp: torch.Tensor, out: torch.Tensor) -> None:
    B.check(B.lib.tf21_tip5_hash_10_dev(_p(inp), inp.numel() // 10, _p(out), _stream()))


def tip5_hash_rows(rows: torch.Tensor, row_len: int, out: torch.Tensor) -> None:
    n_rows = rows.numel() // row_len if row_len else out.numel() // 5
    B.check(B.lib.tf21_tip5_hash_rows_dev(_p(rows), row_len, n_rows, _p(out), _stream()))


def tip5_hash_columns(cols: torch.Tensor, n_rows: int, n_cols: int, out: torch.Tensor, col_stride: int = 0) -> None:
    """digest i = hash_varlen(col_0[i], .., col_{n_cols-1}[i]) over column-major codewords"""
    B.check(B.lib.tf21_tip5_hash_columns_dev(_p(cols), n_rows, n_cols, col_stride or n_rows, _p(out), _stream()))


def merkle_build(leafs: torch.Tensor, nodes: torch.Tensor) -> None:
    B.check(B.lib.tf21_merkle_build_dev(_p(leafs), leafs.numel() // 5, _p(nodes), _stream()))


def merkle_root(leafs: torch.Tensor, root: torch.Tensor) -> None:
    B.check(B.lib.tf21_merkle_root_dev(_p(leafs), leafs.numel() // 5, _p(root), _stream()))


def merkle_scatter_subtree(local_nodes: torch.Tensor, shard: int, n_shards: int, global_nodes: torch.Tensor) -> None:
    B.check(B.lib.tf21_merkle_scatter_subtree_dev(_p(local_nodes), local_nodes.numel() // 10, shard, n_shards,
                                                  _p(global_nodes), _stream()))


def merkle_authentication_structure(nodes: torch.Tensor, leaf_indices, out: torch.Tensor) -> int:
    """out[k] = nodes[needed index k] (MerkleTree::authentication_structure); returns the digest count"""
    import ctypes

    import numpy as np

    idx = np.ascontiguousarray(np.array(leaf_indices, dtype=np.uint64))
    count = ctypes.c_uint64(0)
    B.check(B.lib.tf21_merkle_authentication_structure_dev(_p(nodes), nodes.numel() // 10, idx.ctypes.data if idx.size else None,
                                                           idx.size, _p(out), out.numel() // 5, ctypes.byref(count),
                                                           _stream()))
    return int(count.value)


def mmr_peaks_from_leafs(leafs: torch.Tensor, peaks: torch.Tensor) -> int:
    import ctypes

    n_peaks = ctypes.c_uint64(0)
    B.check(B.lib.tf21_mmr_peaks_from_leafs_dev(_p(leafs), leafs.numel() // 5, _p(peaks), ctypes.byref(n_peaks), _stream()))
    return int(n_peaks.value)


def mmr_bag_peaks(peaks: torch.Tensor, leaf_count: int, out: torch.Tensor) -> None:
    B.check(B.lib.tf21_mmr_bag_peaks_dev(_p(peaks), peaks.numel() // 5, leaf_count, _p(out), _stream()))


def batch_coset_extrapolate(offset_raw: int, codeword_length: int, codewords: torch.Tensor, width: int, points,
                            out: torch.Tensor) -> None:
    """points: numpy uint64 host array of raw words (n_points * width)"""
    import numpy as np

    pts = np.ascontiguousarray(points, dtype=np.uint64)
    n_cw = codewords.numel() // (codeword_length * width)
    B.check(B.lib.tf21_batch_coset_extrapolate_dev(offset_raw, codeword_length, _p(codewords), n_cw, width,
                                                   pts.ctypes.data, pts.size // width, _p(out), _stream()))


def kernel_launch_count() -> int:
    return int(B.lib.tf21_kernel_launch_count())


def profile_enable(on: bool) -> None:
    B.check(B.lib.tf21_profile_enable(int(on)))


def profile_read():
    """[(kernel_name, ms), ...] in launch order"""
    import ctypes

    need = B.lib.tf21_profile_read(None, 0)
    if need < 0:
        B.check(int(need))
    buf = ctypes.create_string_buffer(int(need) + 16)
    got = B.lib.tf21_profile_read(buf, len(buf))
    if got < 0:
        B.check(int(got))
    out = []
    for line in buf.value.decode().splitlines():
        name, ms = line.rsplit(" ", 1)
        out.append((name, float(ms)))
    return out
