"""twenty-first_b200 -- B200-native (sm_100a) drop-in for the STARK hot path of
Neptune-Crypto/twenty-first: Goldilocks NTT/iNTT, coset evaluate/interpolate/LDE, Tip5 and the
Tip5 Merkle build.  The product is the CUDA library `libtf21.so` behind the C ABI of
include/tf21.h; this package is the thin host mirror of the reference's Rust interface used by
the tests and the benchmark (the Rust shim itself is in host/rust/).

Import with importlib (the directory name contains a hyphen):
    tf = importlib.import_module("twenty-first_b200")
"""
from ._binding import (  # noqa: F401
    E_ALLOC,
    E_BAD_ARG,
    E_CAPACITY,
    E_CUDA,
    E_INCORRECT_NUMBER_OF_LEAFS,
    E_LEAF_INDEX_INVALID,
    E_LEN_NOT_POW2,
    E_NCCL,
    E_LEN_TOO_LARGE,
    E_ORDER_LE_DEGREE,
    E_TOO_FEW_LEAFS,
    LIB_PATH,
    SIGNATURES,
    Tf21Error,
    check,
    lib,
)
from .api import (  # noqa: F401
    BFieldElement,
    Digest,
    MerkleTree,
    MerkleTreeError,
    MmrAccumulator,
    Polynomial,
    Tip5,
    intt,
    ntt,
)
from . import device  # noqa: F401
