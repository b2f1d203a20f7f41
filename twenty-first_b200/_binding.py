"""ctypes binding of libtf21.so -- the same C ABI (include/tf21.h) a Rust shim would bind.

There is no fallback: if the CUDA library is missing this module raises at import time.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TF21_LIB") or os.path.join(_HERE, "libtf21.so")  # TF21_LIB: developer A/B builds

OK = 0
E_LEN_NOT_POW2 = -1
E_LEN_TOO_LARGE = -2
E_TOO_FEW_LEAFS = -3
E_INCORRECT_NUMBER_OF_LEAFS = -4
E_ORDER_LE_DEGREE = -5
E_ALLOC = -6
E_CUDA = -7
E_BAD_ARG = -8
E_LEAF_INDEX_INVALID = -9
E_CAPACITY = -10
E_DIVISION_BY_ZERO = -11
E_NCCL = -12

u64 = ctypes.c_uint64
u32 = ctypes.c_uint32
i32 = ctypes.c_int
vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/tf21.h declares
SIGNATURES = {
    "tf21_init": (i32, [i32]),
    "tf21_set_device": (i32, [i32]),
    "tf21_device_count": (i32, []),
    "tf21_sharded_uses_nccl": (i32, [u32]),
    "tf21_shutdown": (i32, []),
    "tf21_strerror": (ctypes.c_char_p, [i32]),
    "tf21_last_cuda_error": (ctypes.c_char_p, []),
    "tf21_kernel_launch_count": (u64, []),
    "tf21_profile_enable": (i32, [i32]),
    "tf21_profile_read": (ctypes.c_int64, [ctypes.c_char_p, u64]),
    "tf21_malloc": (i32, [ctypes.POINTER(vp), u64]),
    "tf21_free": (i32, [vp]),
    "tf21_memcpy_h2d": (i32, [vp, vp, u64, vp]),
    "tf21_memcpy_d2h": (i32, [vp, vp, u64, vp]),
    "tf21_stream_sync": (i32, [vp]),
    "tf21_selftest_field_dev": (i32, [i32, vp, vp, vp, u64, vp]),
    "tf21_selftest_tma_tile_dev": (i32, [vp, u64, u64, vp, vp]),
    "tf21_ntt": (i32, [vp, u64, u32, u64]),
    "tf21_intt": (i32, [vp, u64, u32, u64]),
    "tf21_ntt_dev": (i32, [vp, u64, u32, u64, i32, vp]),
    "tf21_coset_evaluate": (i32, [vp, u64, u32, u64, u64, vp]),
    "tf21_coset_interpolate": (i32, [vp, u64, u32, u64, vp]),
    "tf21_coset_lde": (i32, [vp, u64, u64, u64, u64, u32, vp]),
    "tf21_coset_evaluate_dev": (i32, [vp, u64, u32, u64, u64, vp, vp]),
    "tf21_coset_interpolate_dev": (i32, [vp, u64, u32, u64, vp, vp]),
    "tf21_coset_lde_dev": (i32, [vp, u64, u64, u64, u64, u32, vp, vp]),
    "tf21_poly_mul": (i32, [vp, u64, vp, u64, u32, vp]),
    "tf21_poly_mul_dev": (i32, [vp, u64, vp, u64, u32, vp, vp]),
    "tf21_poly_evaluate_batch_dev": (i32, [vp, u64, u64, u32, vp, u64, vp, vp]),
    "tf21_batch_coset_extrapolate": (i32, [u64, u64, vp, u64, u32, vp, u64, vp]),
    "tf21_batch_coset_extrapolate_dev": (i32, [u64, u64, vp, u64, u32, vp, u64, vp, vp]),
    "tf21_tip5_sample_indices": (i32, [vp, u32, u64, vp]),
    "tf21_poly_reduce_by_ntt_friendly_modulus": (i32, [vp, u64, u32, vp, u64, u64, vp, ctypes.POINTER(u64)]),
    "tf21_poly_clean_divide": (i32, [vp, u64, vp, u64, vp, ctypes.POINTER(u64)]),
    "tf21_poly_square": (i32, [vp, u64, u32, vp]),
    "tf21_poly_square_dev": (i32, [vp, u64, u32, vp, vp]),
    "tf21_tip5_permute": (i32, [vp, u64]),
    "tf21_tip5_hash_10": (i32, [vp, u64, vp]),
    "tf21_tip5_hash_pairs": (i32, [vp, u64, vp]),
    "tf21_tip5_hash_varlen": (i32, [vp, u64, vp]),
    "tf21_tip5_hash_rows": (i32, [vp, u64, u64, vp]),
    "tf21_tip5_permute_dev": (i32, [vp, u64, vp]),
    "tf21_tip5_hash_10_dev": (i32, [vp, u64, vp, vp]),
    "tf21_tip5_hash_rows_dev": (i32, [vp, u64, u64, vp, vp]),
    "tf21_tip5_hash_columns_dev": (i32, [vp, u64, u64, u64, vp, vp]),
    "tf21_merkle_build": (i32, [vp, u64, vp]),
    "tf21_merkle_root": (i32, [vp, u64, vp]),
    "tf21_merkle_build_dev": (i32, [vp, u64, vp, vp]),
    "tf21_merkle_root_dev": (i32, [vp, u64, vp, vp]),
    "tf21_merkle_scatter_subtree_dev": (i32, [vp, u64, u64, u64, vp, vp]),
    "tf21_ntt_sharded": (i32, [vp, u64, u32, u64, i32, u32]),
    "tf21_merkle_build_sharded": (i32, [vp, u64, vp, u32]),
    "tf21_merkle_auth_structure_node_indices": (i32, [u64, vp, u64, vp, u64, ctypes.POINTER(u64)]),
    "tf21_merkle_authentication_structure_dev": (i32, [vp, u64, vp, u64, vp, u64, ctypes.POINTER(u64), vp]),
    "tf21_merkle_authentication_structure_from_leafs": (i32, [vp, u64, vp, u64, vp, u64, ctypes.POINTER(u64)]),
    "tf21_mmr_peaks_from_leafs": (i32, [vp, u64, vp, ctypes.POINTER(u64)]),
    "tf21_mmr_peaks_from_leafs_dev": (i32, [vp, u64, vp, ctypes.POINTER(u64), vp]),
    "tf21_mmr_bag_peaks": (i32, [vp, u64, u64, vp]),
    "tf21_mmr_bag_peaks_dev": (i32, [vp, u64, u64, vp, vp]),
}


def load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class Tf21Error(RuntimeError):
    def __init__(self, code: int):
        self.code = code
        msg = lib.tf21_strerror(code).decode()
        if code == E_CUDA:
            msg += ": " + lib.tf21_last_cuda_error().decode()
        super().__init__(f"tf21 error {code}: {msg}")


def check(code: int) -> None:
    if code != OK:
        raise Tf21Error(code)
