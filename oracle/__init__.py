"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of the plain-C CPU restatement of twenty-first's hot path (oracle/oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (twenty-first_b200) never does.

All arrays are numpy uint64 holding raw Montgomery words, i.e. the in-memory representation of
the reference's BFieldElement (twenty-first/src/math/b_field_element.rs:84-86).

Parity status: pinned against the reference's known-answer tests (tests/test_oracle_kat.py).
The Rust reference itself cannot be built here (no cargo/rustc), hence no oracle/_ref.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

P = 0xFFFFFFFF00000001

E_LEN_NOT_POW2 = -1
E_LEN_TOO_LARGE = -2
E_TOO_FEW_LEAFS = -3
E_INCORRECT_NUMBER_OF_LEAFS = -4
E_ORDER_LE_DEGREE = -5
E_LEAF_INDEX_INVALID = -9


_AVX512_FLAGS = ["-mavx512f", "-mavx512bw", "-mavx512ifma", "-mavx512vbmi"]


def build(native: bool = False, force: bool = False) -> str:
    """Compile oracle.c (+ tip5_avx512.c with the four -mavx512* flags, used only when the HOST has them) with gcc.
    native=True uses -march=native for oracle.c (CPU-baseline timing on the box the benchmark runs on) and writes a
    separate file."""
    out = os.path.join(_BUILD, "liboracle_native.so" if native else "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "field.h", "tip5_mds_generated.h", "tip5_avx512.c")]
    if not force and os.path.exists(out) and all(
        os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs
    ):
        return out
    os.makedirs(_BUILD, exist_ok=True)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    march = "native" if native else "x86-64-v3"
    tag = "native" if native else "base"
    avx_obj = os.path.join(_BUILD, f"tip5_avx512_{tag}.o")
    avx = subprocess.run([cc, "-O3", "-march=x86-64-v3", *_AVX512_FLAGS, "-fPIC", "-std=gnu11", "-c",
                          "-o", avx_obj, os.path.join(_HERE, "tip5_avx512.c")], cwd=_HERE, capture_output=True)
    cmd = [cc, "-O3", f"-march={march}", "-fopenmp", "-fPIC", "-std=gnu11", "-shared", "-o", out,
           os.path.join(_HERE, "oracle.c")]
    if avx.returncode == 0:  # a compiler without the intrinsics just leaves the AVX-512 leg out
        cmd += ["-DTF21_ORACLE_AVX512", avx_obj]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return out


_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


class Oracle:
    def __init__(self, native: bool = False):
        path = build(native=native)
        try:
            self.lib = ctypes.CDLL(path)
        except OSError:
            path = build(native=native, force=True)
            self.lib = ctypes.CDLL(path)
        L = self.lib
        u64, u32, i32 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
        L.oracle_ntt.argtypes = [_u64p, u64, u32]
        L.oracle_intt.argtypes = [_u64p, u64, u32]
        L.oracle_ntt_batch.argtypes = [_u64p, u64, u32, u64, i32, i32]
        L.oracle_poly_scale.argtypes = [_u64p, u64, u32, u64]
        L.oracle_poly_scale.restype = None
        L.oracle_coset_evaluate.argtypes = [_u64p, u64, u32, u64, u64, _u64p]
        L.oracle_coset_interpolate.argtypes = [_u64p, u64, u32, u64, _u64p]
        L.oracle_poly_evaluate.argtypes = [_u64p, u64, u64]
        L.oracle_poly_evaluate.restype = u64
        L.oracle_poly_naive_multiply.argtypes = [_u64p, u64, _u64p, u64, u32, _u64p]
        L.oracle_poly_naive_multiply.restype = None
        L.oracle_poly_fast_multiply.argtypes = [_u64p, u64, _u64p, u64, u32, _u64p]
        L.oracle_tip5_permutation.argtypes = [_u64p]
        L.oracle_tip5_permutation.restype = None
        L.oracle_tip5_hash_rows_batch.argtypes = [_u64p, u64, u64, _u64p, i32]
        L.oracle_tip5_hash_rows_batch.restype = None
        L.oracle_tip5_avx512_available.restype = i32
        L.oracle_tip5_set_impl.argtypes = [i32]
        L.oracle_tip5_set_impl.restype = i32
        L.oracle_tip5_round_avx512.argtypes = [_u64p, i32]
        L.oracle_tip5_round_avx512.restype = i32
        L.oracle_tip5_round.argtypes = [_u64p, i32]
        L.oracle_tip5_round.restype = None
        L.oracle_tip5_round_naive.argtypes = [_u64p, i32]
        L.oracle_tip5_round_naive.restype = None
        L.oracle_tip5_hash_10.argtypes = [_u64p, _u64p]
        L.oracle_tip5_hash_10.restype = None
        L.oracle_tip5_hash_pair.argtypes = [_u64p, _u64p, _u64p]
        L.oracle_tip5_hash_pair.restype = None
        L.oracle_tip5_hash_varlen.argtypes = [_u64p, u64, _u64p]
        L.oracle_tip5_hash_varlen.restype = None
        L.oracle_tip5_hasher_bytes.argtypes = [_u8p, u64]
        L.oracle_tip5_hasher_bytes.restype = u64
        L.oracle_tip5_permute_batch.argtypes = [_u64p, u64, i32]
        L.oracle_tip5_permute_batch.restype = None
        L.oracle_tip5_hash_pairs_batch.argtypes = [_u64p, u64, _u64p, i32]
        L.oracle_tip5_hash_pairs_batch.restype = None
        L.oracle_digest_to_hex.argtypes = [_u64p, ctypes.c_char_p]
        L.oracle_digest_to_hex.restype = None
        L.oracle_merkle_sequential_new.argtypes = [_u64p, u64, _u64p]
        L.oracle_merkle_par_new.argtypes = [_u64p, u64, _u64p, i32, u64]
        L.oracle_merkle_sequential_frugal_root.argtypes = [_u64p, u64, _u64p]
        L.oracle_merkle_par_frugal_root.argtypes = [_u64p, u64, _u64p, i32, u64]
        L.oracle_poly_reduce_by_ntt_friendly_modulus.argtypes = [_u64p, u64, u32, _u64p, u64, u64, _u64p]
        L.oracle_poly_reduce_by_ntt_friendly_modulus.restype = ctypes.c_int64
        L.oracle_poly_naive_divide.argtypes = [_u64p, u64, _u64p, u64, _u64p, _u64p]
        L.oracle_poly_naive_divide.restype = ctypes.c_int64
        L.oracle_poly_evaluate_w.argtypes = [_u64p, u64, u32, _u64p, _u64p]
        L.oracle_poly_evaluate_w.restype = None
        L.oracle_batch_coset_extrapolate.argtypes = [u64, u64, _u64p, u64, u32, _u64p, u64, _u64p]
        L.oracle_tip5_sample_indices.argtypes = [_u64p, u32, u64, ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_tip5_sample_indices.restype = None
        L.oracle_mmr_peaks_from_leafs.argtypes = [_u64p, u64, _u64p]
        L.oracle_mmr_peaks_from_leafs.restype = u64
        L.oracle_mmr_bag_peaks.argtypes = [_u64p, u64, u64, _u64p]
        L.oracle_mmr_bag_peaks.restype = None
        L.oracle_auth_structure_node_indices.argtypes = [u64, _u64p, u64, _u64p]
        L.oracle_auth_structure_node_indices.restype = ctypes.c_int64
        for name in ("new", "value", "inverse_or_zero", "primitive_root_of_unity"):
            f = getattr(L, f"oracle_bfe_{name}")
            f.argtypes = [u64]
            f.restype = u64
        for name in ("add", "sub", "mul", "mod_pow"):
            f = getattr(L, f"oracle_bfe_{name}")
            f.argtypes = [u64, u64]
            f.restype = u64
        L.oracle_bfe_new_array.argtypes = [_u64p, u64]
        L.oracle_bfe_new_array.restype = None
        L.oracle_bfe_value_array.argtypes = [_u64p, u64]
        L.oracle_bfe_value_array.restype = None
        L.oracle_num_threads.restype = i32

    # ---- field ---------------------------------------------------------------------------
    def bfe_new(self, v: int) -> int:
        return self.lib.oracle_bfe_new(v % (1 << 64))

    def bfe_value(self, raw: int) -> int:
        return self.lib.oracle_bfe_value(raw)

    def bfe_add(self, a, b):
        return self.lib.oracle_bfe_add(a, b)

    def bfe_sub(self, a, b):
        return self.lib.oracle_bfe_sub(a, b)

    def bfe_mul(self, a, b):
        return self.lib.oracle_bfe_mul(a, b)

    def bfe_mod_pow(self, a, e):
        return self.lib.oracle_bfe_mod_pow(a, e)

    def bfe_inverse_or_zero(self, a):
        return self.lib.oracle_bfe_inverse_or_zero(a)

    def primitive_root_of_unity(self, n: int) -> int:
        return self.lib.oracle_bfe_primitive_root_of_unity(n)

    def to_raw(self, values) -> np.ndarray:
        """canonical values -> raw Montgomery words (BFieldElement::new element-wise)"""
        a = np.ascontiguousarray(np.array(values, dtype=np.uint64).copy())
        self.lib.oracle_bfe_new_array(_ptr(a.reshape(-1)), a.size)
        return a

    def to_values(self, raw) -> np.ndarray:
        a = np.ascontiguousarray(np.array(raw, dtype=np.uint64).copy())
        self.lib.oracle_bfe_value_array(_ptr(a.reshape(-1)), a.size)
        return a

    # ---- ntt -----------------------------------------------------------------------------
    def ntt(self, x: np.ndarray, width: int = 1) -> int:
        """in place; x has n*width words"""
        return self.lib.oracle_ntt(_ptr(x), x.size // width, width)

    def intt(self, x: np.ndarray, width: int = 1) -> int:
        return self.lib.oracle_intt(_ptr(x), x.size // width, width)

    def ntt_batch(self, x: np.ndarray, n: int, width: int, batch: int, inverse: bool,
                  threads: int = 0) -> int:
        assert x.size == n * width * batch
        return self.lib.oracle_ntt_batch(_ptr(x), n, width, batch, int(inverse), threads)

    # ---- polynomial ------------------------------------------------------------------------
    def poly_scale(self, coeffs: np.ndarray, width: int, alpha_raw: int) -> None:
        self.lib.oracle_poly_scale(_ptr(coeffs), coeffs.size // width, width, alpha_raw)

    def coset_evaluate(self, coeffs: np.ndarray, width: int, offset_raw: int, order: int):
        out = np.zeros(order * width, dtype=np.uint64)
        rc = self.lib.oracle_coset_evaluate(_ptr(coeffs), coeffs.size // width, width,
                                            offset_raw, order, _ptr(out))
        return rc, out

    def coset_interpolate(self, values: np.ndarray, width: int, offset_raw: int):
        out = np.zeros(values.size, dtype=np.uint64)
        rc = self.lib.oracle_coset_interpolate(_ptr(values), values.size // width, width,
                                               offset_raw, _ptr(out))
        return rc, out

    def poly_naive_multiply(self, a: np.ndarray, b: np.ndarray, width: int) -> np.ndarray:
        na, nb = a.size // width, b.size // width
        out = np.zeros(max(0, na + nb - 1) * width if na and nb else 0, dtype=np.uint64)
        if out.size:
            self.lib.oracle_poly_naive_multiply(_ptr(a), na, _ptr(b), nb, width, _ptr(out))
        return out

    def poly_fast_multiply(self, a: np.ndarray, b: np.ndarray, width: int):
        na, nb = a.size // width, b.size // width
        out = np.zeros(max(0, na + nb - 1) * width if na and nb else 0, dtype=np.uint64)
        rc = self.lib.oracle_poly_fast_multiply(_ptr(a), na, _ptr(b), nb, width, _ptr(out)) if out.size else 0
        return rc, out

    def poly_evaluate(self, coeffs: np.ndarray, x_raw: int) -> int:
        return self.lib.oracle_poly_evaluate(_ptr(coeffs), coeffs.size, x_raw)

    # ---- tip5 ------------------------------------------------------------------------------
    def tip5_permutation(self, state: np.ndarray) -> None:
        assert state.size == 16
        self.lib.oracle_tip5_permutation(_ptr(state))

    def tip5_round(self, state: np.ndarray, round_index: int, naive: bool = False) -> None:
        """one round in place: the scalar build's form (mds_generated) or NaiveTip5's (tip5/naive.rs:26-76)"""
        assert state.size == 16
        (self.lib.oracle_tip5_round_naive if naive else self.lib.oracle_tip5_round)(_ptr(state), round_index)

    def tip5_avx512_available(self) -> bool:
        """the AVX-512 IFMA/VBMI round (tip5/avx512.rs) was built and this host can run it"""
        return bool(self.lib.oracle_tip5_avx512_available())

    def tip5_set_impl(self, impl: str) -> str:
        """'scalar' (mds_generated, tip5/mod.rs:175-506) or 'avx512' (tip5/avx512.rs); returns what is in effect"""
        return "avx512" if self.lib.oracle_tip5_set_impl(1 if impl == "avx512" else 0) == 1 else "scalar"

    def tip5_round_avx512(self, state: np.ndarray, round_index: int) -> bool:
        assert state.size == 16
        return self.lib.oracle_tip5_round_avx512(_ptr(state), round_index) == 0

    def tip5_permute_batch(self, states: np.ndarray, threads: int = 0) -> None:
        self.lib.oracle_tip5_permute_batch(_ptr(states), states.size // 16, threads)

    def hash_10(self, inp: np.ndarray) -> np.ndarray:
        out = np.zeros(5, dtype=np.uint64)
        self.lib.oracle_tip5_hash_10(_ptr(inp), _ptr(out))
        return out

    def hash_pair(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        out = np.zeros(5, dtype=np.uint64)
        self.lib.oracle_tip5_hash_pair(_ptr(left), _ptr(right), _ptr(out))
        return out

    def hash_pairs_batch(self, pairs: np.ndarray, threads: int = 0) -> np.ndarray:
        count = pairs.size // 10
        out = np.zeros(5 * count, dtype=np.uint64)
        self.lib.oracle_tip5_hash_pairs_batch(_ptr(pairs), count, _ptr(out), threads)
        return out

    def hash_varlen(self, inp: np.ndarray) -> np.ndarray:
        out = np.zeros(5, dtype=np.uint64)
        inp = np.ascontiguousarray(inp, dtype=np.uint64)
        buf = inp if inp.size else np.zeros(1, dtype=np.uint64)
        self.lib.oracle_tip5_hash_varlen(_ptr(buf), inp.size, _ptr(out))
        return out

    def hash_rows(self, rows: np.ndarray, threads: int = 0) -> np.ndarray:
        """hash_varlen of every row of a (n_rows, row_len) matrix -> (n_rows, 5)"""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        n_rows, row_len = rows.shape
        out = np.zeros((n_rows, 5), dtype=np.uint64)
        src = rows if rows.size else np.zeros(1, dtype=np.uint64)
        self.lib.oracle_tip5_hash_rows_batch(_ptr(src.reshape(-1)), row_len, n_rows, _ptr(out.reshape(-1)), threads)
        return out

    def hasher_bytes(self, data: bytes) -> int:
        buf = (ctypes.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
        return self.lib.oracle_tip5_hasher_bytes(buf, len(data))

    def digest_to_hex(self, digest_raw: np.ndarray) -> str:
        out = ctypes.create_string_buffer(81)
        self.lib.oracle_digest_to_hex(_ptr(np.ascontiguousarray(digest_raw, dtype=np.uint64)), out)
        return out.value.decode()

    # ---- merkle ----------------------------------------------------------------------------
    def merkle_sequential_new(self, leafs: np.ndarray):
        n = leafs.size // 5
        nodes = np.zeros(max(1, 10 * n), dtype=np.uint64)
        rc = self.lib.oracle_merkle_sequential_new(_ptr(leafs if n else nodes), n, _ptr(nodes))
        return rc, nodes[: 10 * n]

    def merkle_par_new(self, leafs: np.ndarray, threads: int = 0, cutoff: int = 512):
        n = leafs.size // 5
        nodes = np.zeros(max(1, 10 * n), dtype=np.uint64)
        rc = self.lib.oracle_merkle_par_new(_ptr(leafs if n else nodes), n, _ptr(nodes), threads, cutoff)
        return rc, nodes[: 10 * n]

    def merkle_sequential_frugal_root(self, leafs: np.ndarray):
        n = leafs.size // 5
        root = np.zeros(5, dtype=np.uint64)
        rc = self.lib.oracle_merkle_sequential_frugal_root(_ptr(leafs if n else root), n, _ptr(root))
        return rc, root

    def merkle_par_frugal_root(self, leafs: np.ndarray, threads: int = 0, cutoff: int = 512):
        n = leafs.size // 5
        root = np.zeros(5, dtype=np.uint64)
        rc = self.lib.oracle_merkle_par_frugal_root(_ptr(leafs if n else root), n, _ptr(root), threads, cutoff)
        return rc, root

    def poly_reduce_by_ntt_friendly_modulus(self, coeffs: np.ndarray, width: int, shift_ntt: np.ndarray, tail_length: int):
        n, dl = coeffs.size // width, shift_ntt.size // width
        out = np.zeros(max(1, max(n, dl) * width), dtype=np.uint64)
        buf = coeffs if coeffs.size else np.zeros(width, dtype=np.uint64)
        k = self.lib.oracle_poly_reduce_by_ntt_friendly_modulus(_ptr(buf), n, width, _ptr(shift_ntt), dl, tail_length,
                                                                _ptr(out))
        return k, out[: max(0, k) * width].copy()

    def poly_naive_divide(self, a: np.ndarray, b: np.ndarray):
        """(quotient, remainder) of Polynomial::naive_divide over BFieldElement; None for a zero divisor"""
        quot = np.zeros(max(1, a.size), dtype=np.uint64)
        rem = np.zeros(max(1, a.size), dtype=np.uint64)
        k = self.lib.oracle_poly_naive_divide(_ptr(a if a.size else rem), a.size, _ptr(b if b.size else rem), b.size,
                                              _ptr(quot), _ptr(rem))
        if k < 0:
            return None
        return quot[:k].copy(), rem[: a.size].copy()

    def poly_evaluate_w(self, coeffs: np.ndarray, width: int, x: np.ndarray) -> np.ndarray:
        out = np.zeros(width, dtype=np.uint64)
        buf = coeffs if coeffs.size else np.zeros(width, dtype=np.uint64)
        self.lib.oracle_poly_evaluate_w(_ptr(buf), coeffs.size // width, width, _ptr(np.ascontiguousarray(x)), _ptr(out))
        return out

    def batch_coset_extrapolate(self, offset_raw: int, n: int, codewords: np.ndarray, width: int, points: np.ndarray):
        n_cw = codewords.size // (n * width)
        n_pts = points.size // width
        out = np.zeros(max(1, n_cw * n_pts * width), dtype=np.uint64)
        rc = self.lib.oracle_batch_coset_extrapolate(offset_raw, n, _ptr(codewords), n_cw, width, _ptr(points), n_pts,
                                                     _ptr(out))
        return rc, out[: n_cw * n_pts * width]

    def tip5_sample_indices(self, state: np.ndarray, upper_bound: int, num_indices: int) -> np.ndarray:
        """state (16 raw words) is advanced in place like `&mut self`"""
        out = np.zeros(max(1, num_indices), dtype=np.uint32)
        self.lib.oracle_tip5_sample_indices(_ptr(state), upper_bound, num_indices,
                                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        return out[:num_indices]

    def mmr_peaks_from_leafs(self, leafs: np.ndarray) -> np.ndarray:
        """MmrAccumulator::peaks_from_leafs (mmr_accumulator.rs:96-115), any leaf count"""
        n = leafs.size // 5
        peaks = np.zeros(64 * 5, dtype=np.uint64)
        k = self.lib.oracle_mmr_peaks_from_leafs(_ptr(leafs if n else peaks), n, _ptr(peaks))
        return peaks[: 5 * k].reshape(-1, 5).copy()

    def mmr_bag_peaks(self, peaks: np.ndarray, leaf_count: int) -> np.ndarray:
        out = np.zeros(5, dtype=np.uint64)
        peaks = np.ascontiguousarray(peaks, dtype=np.uint64).reshape(-1)
        buf = peaks if peaks.size else np.zeros(5, dtype=np.uint64)
        self.lib.oracle_mmr_bag_peaks(_ptr(buf), peaks.size // 5, leaf_count, _ptr(out))
        return out

    def auth_structure_node_indices(self, num_leafs: int, leaf_indices):
        """(rc, indices): MerkleTree::authentication_structure_node_indices (merkle_tree.rs:449-504)"""
        idx = np.ascontiguousarray(np.array(leaf_indices, dtype=np.uint64))
        height = max(1, int(num_leafs).bit_length())
        out = np.zeros(idx.size * height + 1, dtype=np.uint64)
        buf = idx if idx.size else np.zeros(1, dtype=np.uint64)
        rc = self.lib.oracle_auth_structure_node_indices(num_leafs, _ptr(buf), idx.size, _ptr(out))
        if rc < 0:
            return rc, None
        return 0, out[:rc].copy()

    def num_threads(self) -> int:
        return self.lib.oracle_num_threads()

    def set_num_threads(self, n: int) -> None:
        self.lib.oracle_set_num_threads(int(n))


_default = None


def get(native: bool = False) -> Oracle:
    global _default
    if native:
        return Oracle(native=True)
    if _default is None:
        _default = Oracle()
    return _default


def splitmix64_words(seed: int, count: int) -> np.ndarray:
    """SURVEY.md 8(d) synthetic-input generator: SplitMix64 stream, values >= p rejected, used
    directly as raw words (x R is a bijection on F_p).  Vectorised: draws with a small surplus and
    filters (deterministic for a given seed and count)."""
    out = np.empty(count, dtype=np.uint64)
    filled = 0
    state = np.uint64(seed)
    GAMMA = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        while filled < count:
            m = max(1024, int((count - filled) * 1.001) + 16)
            idx = np.arange(1, m + 1, dtype=np.uint64)
            z = state + idx * GAMMA
            state = z[-1]
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            z = z[z < np.uint64(P)]
            take = min(z.size, count - filled)
            out[filled:filled + take] = z[:take]
            filled += take
    return out
