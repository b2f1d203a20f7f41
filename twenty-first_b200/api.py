"""Host mirror of the reference's public interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the Rust crate:
  math::ntt::{ntt, intt}                          twenty-first/src/math/ntt.rs:67,109
  Polynomial::{scale-free} fast_coset_evaluate    math/polynomial.rs:1374-1399
  Polynomial::fast_coset_interpolate              math/polynomial.rs:1907-1918
  Tip5::{permutation, hash_10, hash_pair, hash_varlen}   tip5/mod.rs:529-623
  MerkleTree::{par_new, sequential_new, par_frugal_root, sequential_frugal_root, root, node,
               leafs, num_leafs, height}          util_types/merkle_tree.rs:149-364, 624-653
  MerkleTree::{authentication_structure_node_indices, authentication_structure,
               par_authentication_structure_from_leafs}   util_types/merkle_tree.rs:449-542, 614-622
  MmrAccumulator::{new_from_leafs, peaks, bag_peaks}     util_types/mmr/mmr_accumulator.rs:29-34, 96-134, 379-391

Arrays are numpy uint64 of *raw Montgomery words* -- exactly the bytes of a Rust
`&[BFieldElement]` / `&[XFieldElement]` (width 3) / `&[Digest]` (5 words).  Where the reference
panics this raises `AssertionError`-like `Tf21Error` with the same message; where it returns
`Err(MerkleTreeError::..)` this raises `MerkleTreeError`.
"""
from __future__ import annotations

import numpy as np

from . import _binding as B

P = 0xFFFFFFFF00000001
_R = (1 << 64) % P
_R_INV = pow(_R, P - 2, P)


def _words(a: np.ndarray) -> np.ndarray:
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]):
        raise TypeError("expected a C-contiguous numpy uint64 array of raw words")
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


class BFieldElement:
    """Only what the hot path needs: conversion between canonical values and raw words
    (b_field_element.rs:235-237 `new`, :248 `value`)."""

    P = P

    @staticmethod
    def new(value: int) -> int:
        return (value % P) * _R % P

    @staticmethod
    def value(raw: int) -> int:
        return raw % P * _R_INV % P

    @staticmethod
    def generator() -> int:  # b_field_element.rs:312
        return BFieldElement.new(7)


def _width_of(x: np.ndarray) -> int:
    if x.ndim == 1:
        return 1
    if x.ndim == 2 and x.shape[1] == 3:
        return 3
    raise TypeError("expected shape (n,) for BFieldElement or (n, 3) for XFieldElement slices")


def ntt(x: np.ndarray) -> None:
    """In-place NTT of a `[BFieldElement]` (shape (n,)) or `[XFieldElement]` (shape (n,3)) slice.
    Panics (raises) if the length is not a power of two or exceeds u32::MAX (ntt.rs:135-137)."""
    _words(x)
    B.check(B.lib.tf21_ntt(_ptr(x), x.shape[0], _width_of(x), 1))


def intt(x: np.ndarray) -> None:
    _words(x)
    B.check(B.lib.tf21_intt(_ptr(x), x.shape[0], _width_of(x), 1))


def ntt_batch(x: np.ndarray, n: int, width: int = 1, inverse: bool = False) -> None:
    """`batch` contiguous arrays of n*width words, the caller-side `par_iter` over columns."""
    _words(x)
    batch = x.size // (n * width) if n else 0
    fn = B.lib.tf21_intt if inverse else B.lib.tf21_ntt
    B.check(fn(_ptr(x), n, width, batch))


class Polynomial:
    """`Polynomial<FF>` restricted to the coset fast paths."""

    def __init__(self, coefficients: np.ndarray):
        self.coefficients = _words(np.ascontiguousarray(coefficients))
        self.width = _width_of(self.coefficients)

    def fast_coset_evaluate(self, offset_raw: int, order: int) -> np.ndarray:
        """polynomial.rs:1374-1399; panics if `order` <= degree."""
        shape = (order,) if self.width == 1 else (order, 3)
        out = np.zeros(shape, dtype=np.uint64)
        B.check(B.lib.tf21_coset_evaluate(_ptr(self.coefficients), self.coefficients.shape[0], self.width,
                                          offset_raw, order, _ptr(out)))
        return out

    def fast_multiply(self, other: "Polynomial") -> "Polynomial":
        """polynomial.rs:900-932 (same-width operands); trailing zero coefficients are kept"""
        if self.width != other.width:
            raise TypeError("mixed BFieldElement / XFieldElement products are not offloaded")
        na, nb = self.coefficients.shape[0], other.coefficients.shape[0]
        n = na + nb - 1 if na and nb else 0
        out = np.zeros((n,) if self.width == 1 else (n, 3), dtype=np.uint64)
        B.check(B.lib.tf21_poly_mul(_ptr(self.coefficients), na, _ptr(other.coefficients), nb, self.width,
                                    _ptr(out)))
        return Polynomial(out)

    def reduce_by_ntt_friendly_modulus(self, shift_ntt: np.ndarray, tail_length: int) -> "Polynomial":
        """polynomial.rs:1087-1148; panics (raises) unless len(shift_ntt) is a power of two"""
        import ctypes

        shift_ntt = _words(np.ascontiguousarray(shift_ntt))
        n, dl = self.coefficients.shape[0], shift_ntt.shape[0]
        m = min(n, dl)
        out = np.zeros((max(m, 1),) if self.width == 1 else (max(m, 1), 3), dtype=np.uint64)
        n_out = ctypes.c_uint64(0)
        B.check(B.lib.tf21_poly_reduce_by_ntt_friendly_modulus(_ptr(self.coefficients), n, self.width, _ptr(shift_ntt), dl,
                                                               tail_length, _ptr(out), ctypes.byref(n_out)))
        return Polynomial(out[: n_out.value].copy())

    def clean_divide(self, divisor: "Polynomial") -> "Polynomial":
        """polynomial.rs:2358-2413 (BFieldElement only): exact quotient of a division without remainder; panics
        (raises) for a zero divisor"""
        import ctypes

        if self.width != 1 or divisor.width != 1:
            raise TypeError("clean_divide is defined for Polynomial<BFieldElement>")
        na, nb = self.coefficients.shape[0], divisor.coefficients.shape[0]
        out = np.zeros(max(na, 1), dtype=np.uint64)
        n_q = ctypes.c_uint64(0)
        B.check(B.lib.tf21_poly_clean_divide(_ptr(self.coefficients), na, _ptr(divisor.coefficients), nb, _ptr(out),
                                             ctypes.byref(n_q)))
        return Polynomial(out[: n_q.value].copy())

    def fast_square(self) -> "Polynomial":
        """polynomial.rs:780-802; trailing zero coefficients are kept"""
        na = self.coefficients.shape[0]
        n = 2 * na - 1 if na else 0
        out = np.zeros((n,) if self.width == 1 else (n, 3), dtype=np.uint64)
        B.check(B.lib.tf21_poly_square(_ptr(self.coefficients), na, self.width, _ptr(out)))
        return Polynomial(out)

    @staticmethod
    def par_batch_coset_extrapolate(domain_offset_raw: int, codeword_length: int, codewords: np.ndarray,
                                    points: np.ndarray) -> np.ndarray:
        """polynomial.rs:2188-2331 (batch_ / par_batch_coset_extrapolate): codewords and points of the same field;
        result[(codeword, point)] flattened like the reference's flat_map.  Panics (raises) if the codeword
        length is not a power of two."""
        codewords = _words(np.ascontiguousarray(codewords))
        points = _words(np.ascontiguousarray(points))
        w = 3 if (codewords.ndim == 2 and codewords.shape[1] == 3) else 1
        n_cw = codewords.size // (codeword_length * w) if codeword_length else 0
        n_pts = points.size // w
        out = np.zeros((n_cw * n_pts,) if w == 1 else (n_cw * n_pts, 3), dtype=np.uint64)
        B.check(B.lib.tf21_batch_coset_extrapolate(domain_offset_raw, codeword_length, _ptr(codewords), n_cw, w,
                                                   _ptr(points), n_pts, _ptr(out)))
        return out

    batch_coset_extrapolate = par_batch_coset_extrapolate

    @staticmethod
    def coset_extrapolate(domain_offset_raw: int, codeword: np.ndarray, points: np.ndarray) -> np.ndarray:
        """polynomial.rs:2117-2127"""
        return Polynomial.par_batch_coset_extrapolate(domain_offset_raw, codeword.shape[0], codeword, points)

    @staticmethod
    def fast_coset_interpolate(offset_raw: int, values: np.ndarray) -> "Polynomial":
        """polynomial.rs:1907-1918"""
        values = _words(np.ascontiguousarray(values))
        out = np.zeros_like(values)
        B.check(B.lib.tf21_coset_interpolate(_ptr(values), values.shape[0], _width_of(values), offset_raw,
                                             _ptr(out)))
        return Polynomial(out)

    @staticmethod
    def coset_lde(values: np.ndarray, offset_in_raw: int, n_out: int, offset_out_raw: int) -> np.ndarray:
        """fast_coset_interpolate(offset_in, values).fast_coset_evaluate(offset_out, n_out), fused."""
        values = _words(np.ascontiguousarray(values))
        w = _width_of(values)
        out = np.zeros((n_out,) if w == 1 else (n_out, 3), dtype=np.uint64)
        B.check(B.lib.tf21_coset_lde(_ptr(values), values.shape[0], offset_in_raw, n_out, offset_out_raw, w,
                                     _ptr(out)))
        return out


class Digest:
    LEN = 5  # tip5/digest.rs:49

    @staticmethod
    def to_hex(raw: np.ndarray) -> str:
        """digest.rs:85-90,144-152: canonical values, little-endian bytes, lower-case hex"""
        return b"".join(BFieldElement.value(int(v)).to_bytes(8, "little") for v in raw).hex()


class Tip5:
    """Batched forms of Tip5's associated functions; a leading batch axis replaces the caller's
    `par_iter().map(..)` (benches/tip5.rs:43-49)."""

    @staticmethod
    def permutation(states: np.ndarray) -> None:
        """in place; shape (count, 16) or (16,)"""
        _words(states)
        B.check(B.lib.tf21_tip5_permute(_ptr(states), states.size // 16))

    @staticmethod
    def hash_10(inputs: np.ndarray) -> np.ndarray:
        _words(inputs)
        count = inputs.size // 10
        out = np.zeros((count, 5) if inputs.ndim == 2 else 5, dtype=np.uint64)
        B.check(B.lib.tf21_tip5_hash_10(_ptr(inputs), count, _ptr(out)))
        return out

    @staticmethod
    def hash_pair(left: np.ndarray, right: np.ndarray) -> np.ndarray:
        pairs = np.ascontiguousarray(np.concatenate([left.reshape(-1, 5), right.reshape(-1, 5)], axis=1))
        out = np.zeros((pairs.shape[0], 5), dtype=np.uint64)
        B.check(B.lib.tf21_tip5_hash_pairs(_ptr(pairs), pairs.shape[0], _ptr(out)))
        return out.reshape(left.shape)

    @staticmethod
    def hash_varlen(inp: np.ndarray) -> np.ndarray:
        inp = _words(np.ascontiguousarray(inp))
        out = np.zeros(5, dtype=np.uint64)
        B.check(B.lib.tf21_tip5_hash_varlen(_ptr(inp), inp.size, _ptr(out)))
        return out

    @staticmethod
    def sample_indices(state: np.ndarray, upper_bound: int, num_indices: int) -> np.ndarray:
        """tip5/mod.rs:636-656 on the sponge state `state` (16 raw words, advanced in place like `&mut self`);
        panics (raises) unless upper_bound is a power of two"""
        _words(state)
        out = np.zeros(num_indices, dtype=np.uint32)
        B.check(B.lib.tf21_tip5_sample_indices(_ptr(state), upper_bound, num_indices,
                                               out.ctypes.data if num_indices else None))
        return out

    @staticmethod
    def hash_rows(rows: np.ndarray) -> np.ndarray:
        """hash_varlen of every row of a (n_rows, row_len) matrix"""
        rows = _words(np.ascontiguousarray(rows))
        out = np.zeros((rows.shape[0], 5), dtype=np.uint64)
        B.check(B.lib.tf21_tip5_hash_rows(_ptr(rows), rows.shape[1], rows.shape[0], _ptr(out)))
        return out


class MerkleTreeError(Exception):
    """util_types/merkle_tree.rs:933-965"""

    TOO_FEW_LEAFS = "TooFewLeafs"
    INCORRECT_NUMBER_OF_LEAFS = "IncorrectNumberOfLeafs"
    TREE_TOO_HIGH = "TreeTooHigh"
    LEAF_INDEX_INVALID = "LeafIndexInvalid"

    def __init__(self, kind: str):
        self.kind = kind
        super().__init__(kind)


def _merkle_check(code: int) -> None:
    if code == B.E_TOO_FEW_LEAFS:
        raise MerkleTreeError(MerkleTreeError.TOO_FEW_LEAFS)
    if code == B.E_INCORRECT_NUMBER_OF_LEAFS:
        raise MerkleTreeError(MerkleTreeError.INCORRECT_NUMBER_OF_LEAFS)
    if code == B.E_ALLOC:
        raise MerkleTreeError(MerkleTreeError.TREE_TOO_HIGH)
    if code == B.E_LEAF_INDEX_INVALID:
        raise MerkleTreeError(MerkleTreeError.LEAF_INDEX_INVALID)
    B.check(code)


class MerkleTree:
    """Heap-indexed node array exactly like the reference's `Vec<Digest>` (merkle_tree.rs:85-88)."""

    def __init__(self, nodes: np.ndarray):
        self.nodes = nodes  # shape (2n, 5)

    @staticmethod
    def par_new(leafs: np.ndarray) -> "MerkleTree":
        leafs = _words(np.ascontiguousarray(leafs)).reshape(-1, 5)
        n = leafs.shape[0]
        nodes = np.zeros((2 * n, 5), dtype=np.uint64)
        _merkle_check(B.lib.tf21_merkle_build(_ptr(leafs), n, _ptr(nodes)))
        return MerkleTree(nodes)

    sequential_new = par_new  # same result by definition (merkle_tree.rs:1059-1074)

    @staticmethod
    def par_frugal_root(leafs: np.ndarray) -> np.ndarray:
        leafs = _words(np.ascontiguousarray(leafs)).reshape(-1, 5)
        n = leafs.shape[0]
        if n == 0 or n & (n - 1):
            # par_frugal_root checks the power of two first (merkle_tree.rs:333-335)
            raise MerkleTreeError(MerkleTreeError.INCORRECT_NUMBER_OF_LEAFS)
        root = np.zeros(5, dtype=np.uint64)
        _merkle_check(B.lib.tf21_merkle_root(_ptr(leafs), n, _ptr(root)))
        return root

    @staticmethod
    def sequential_frugal_root(leafs: np.ndarray) -> np.ndarray:
        leafs = _words(np.ascontiguousarray(leafs)).reshape(-1, 5)
        n = leafs.shape[0]
        root = np.zeros(5, dtype=np.uint64)
        _merkle_check(B.lib.tf21_merkle_root(_ptr(leafs), n, _ptr(root)))  # empty -> TooFewLeafs (:300-302)
        return root

    def root(self) -> np.ndarray:  # merkle_tree.rs:624
        return self.nodes[1]

    def num_leafs(self) -> int:  # :628
        return self.nodes.shape[0] // 2

    def height(self) -> int:  # :634
        return self.num_leafs().bit_length() - 1

    def node(self, index: int):  # :644
        return self.nodes[index] if 0 <= index < self.nodes.shape[0] else None

    def leafs(self) -> np.ndarray:  # :653
        return self.nodes[self.num_leafs():]

    # ---- authentication structures ----------------------------------------
    @staticmethod
    def authentication_structure_node_indices(num_leafs: int, leaf_indices) -> np.ndarray:
        """merkle_tree.rs:449-504: descending node indices; Err(IncorrectNumberOfLeafs / LeafIndexInvalid)"""
        import ctypes

        idx = np.ascontiguousarray(np.array(leaf_indices, dtype=np.uint64))
        count = ctypes.c_uint64(0)
        rc = B.lib.tf21_merkle_auth_structure_node_indices(num_leafs, _ptr(idx), idx.size, None, 0,
                                                           ctypes.byref(count))
        if rc != B.E_CAPACITY:
            _merkle_check(rc)
        out = np.zeros(count.value, dtype=np.uint64)
        _merkle_check(B.lib.tf21_merkle_auth_structure_node_indices(num_leafs, _ptr(idx), idx.size, _ptr(out),
                                                                    out.size, ctypes.byref(count)))
        return out

    def authentication_structure(self, leaf_indices) -> np.ndarray:
        """merkle_tree.rs:614-622 on a host-resident node array: a plain gather"""
        idx = MerkleTree.authentication_structure_node_indices(self.num_leafs(), leaf_indices)
        return self.nodes[idx.astype(np.int64)]

    @staticmethod
    def par_authentication_structure_from_leafs(leafs: np.ndarray, leaf_indices) -> np.ndarray:
        """merkle_tree.rs:532-542 (and the sequential form :514-523): tree built on the device, gathered there"""
        import ctypes

        leafs = _words(np.ascontiguousarray(leafs)).reshape(-1, 5)
        idx = np.ascontiguousarray(np.array(leaf_indices, dtype=np.uint64))
        need = MerkleTree.authentication_structure_node_indices(leafs.shape[0], idx).size
        out = np.zeros((need, 5), dtype=np.uint64)
        count = ctypes.c_uint64(0)
        _merkle_check(B.lib.tf21_merkle_authentication_structure_from_leafs(
            _ptr(leafs), leafs.shape[0], _ptr(idx), idx.size, _ptr(out), need, ctypes.byref(count)))
        return out

    sequential_authentication_structure_from_leafs = par_authentication_structure_from_leafs


class MmrAccumulator:
    """`MmrAccumulator` restricted to bulk construction and the commitment."""

    def __init__(self, peaks: np.ndarray, leaf_count: int):  # `init`, mmr_accumulator.rs:25-27
        self._peaks = np.ascontiguousarray(peaks, dtype=np.uint64).reshape(-1, 5)
        self.leaf_count = int(leaf_count)

    @staticmethod
    def new_from_leafs(leafs: np.ndarray) -> "MmrAccumulator":
        """mmr_accumulator.rs:29-34 / peaks_from_leafs :96-115; any leaf count including 0"""
        import ctypes

        leafs = _words(np.ascontiguousarray(leafs, dtype=np.uint64)).reshape(-1, 5)
        peaks = np.zeros((64, 5), dtype=np.uint64)
        n_peaks = ctypes.c_uint64(0)
        B.check(B.lib.tf21_mmr_peaks_from_leafs(_ptr(leafs), leafs.shape[0], _ptr(peaks), ctypes.byref(n_peaks)))
        return MmrAccumulator(peaks[: n_peaks.value].copy(), leafs.shape[0])

    def peaks(self) -> np.ndarray:  # :133-135
        return self._peaks

    def num_leafs(self) -> int:  # :143-145
        return self.leaf_count

    def is_consistent(self) -> bool:  # :119-122
        return self._peaks.shape[0] == bin(self.leaf_count).count("1")

    def bag_peaks(self) -> np.ndarray:  # :127-129, 379-391
        out = np.zeros(5, dtype=np.uint64)
        B.check(B.lib.tf21_mmr_bag_peaks(_ptr(self._peaks), self._peaks.shape[0], self.leaf_count, _ptr(out)))
        return out
