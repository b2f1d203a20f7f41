//! Raw bindings of `include/tf21.h`.  One `extern "C"` item per exported symbol, same order as
//! the header.  NOTE: written without a Rust toolchain in the build container (no cargo/rustc
//! there); the ABI itself is exercised through Python ctypes (`twenty-first_b200/_binding.py`),
//! which declares exactly the same signatures.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const TF21_OK: c_int = 0;
pub const TF21_E_LEN_NOT_POW2: c_int = -1;
pub const TF21_E_LEN_TOO_LARGE: c_int = -2;
pub const TF21_E_TOO_FEW_LEAFS: c_int = -3;
pub const TF21_E_INCORRECT_NUMBER_OF_LEAFS: c_int = -4;
pub const TF21_E_ORDER_LE_DEGREE: c_int = -5;
pub const TF21_E_ALLOC: c_int = -6;
pub const TF21_E_CUDA: c_int = -7;
pub const TF21_E_BAD_ARG: c_int = -8;
pub const TF21_E_LEAF_INDEX_INVALID: c_int = -9;
pub const TF21_E_CAPACITY: c_int = -10;
pub const TF21_E_DIVISION_BY_ZERO: c_int = -11;

pub type tf21_stream_t = *mut c_void; // cudaStream_t

extern "C" {
    pub fn tf21_init(device: c_int) -> c_int;
    pub fn tf21_shutdown() -> c_int;
    pub fn tf21_strerror(code: c_int) -> *const c_char;
    pub fn tf21_last_cuda_error() -> *const c_char;
    pub fn tf21_kernel_launch_count() -> u64;
    pub fn tf21_profile_enable(on: c_int) -> c_int;
    pub fn tf21_profile_read(buf: *mut c_char, buflen: u64) -> i64;
    pub fn tf21_malloc(dptr: *mut *mut c_void, bytes: u64) -> c_int;
    pub fn tf21_free(dptr: *mut c_void) -> c_int;
    pub fn tf21_memcpy_h2d(dst: *mut c_void, src: *const c_void, bytes: u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_memcpy_d2h(dst: *mut c_void, src: *const c_void, bytes: u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_stream_sync(s: tf21_stream_t) -> c_int;

    pub fn tf21_ntt(data: *mut u64, n: u64, width: u32, batch: u64) -> c_int;
    pub fn tf21_intt(data: *mut u64, n: u64, width: u32, batch: u64) -> c_int;
    pub fn tf21_ntt_dev(d: *mut u64, n: u64, width: u32, batch: u64, inverse: c_int, s: tf21_stream_t) -> c_int;

    pub fn tf21_coset_evaluate(coeffs: *const u64, n_coeffs: u64, width: u32, offset_raw: u64, order: u64,
                               out: *mut u64) -> c_int;
    pub fn tf21_coset_interpolate(values: *const u64, n: u64, width: u32, offset_raw: u64,
                                  coeffs_out: *mut u64) -> c_int;
    pub fn tf21_coset_lde(values: *const u64, n_in: u64, offset_in_raw: u64, n_out: u64, offset_out_raw: u64,
                          width: u32, out: *mut u64) -> c_int;
    pub fn tf21_coset_evaluate_dev(c: *const u64, n_coeffs: u64, width: u32, offset_raw: u64, order: u64,
                                   out: *mut u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_coset_interpolate_dev(v: *const u64, n: u64, width: u32, offset_raw: u64, out: *mut u64,
                                      s: tf21_stream_t) -> c_int;
    pub fn tf21_coset_lde_dev(v: *const u64, n_in: u64, offset_in_raw: u64, n_out: u64, offset_out_raw: u64,
                              width: u32, out: *mut u64, s: tf21_stream_t) -> c_int;

    pub fn tf21_poly_mul(a: *const u64, n_a: u64, b: *const u64, n_b: u64, width: u32, out: *mut u64) -> c_int;
    pub fn tf21_poly_mul_dev(a: *const u64, n_a: u64, b: *const u64, n_b: u64, width: u32, out: *mut u64,
                             s: tf21_stream_t) -> c_int;

    pub fn tf21_poly_evaluate_batch_dev(polys: *const u64, n: u64, n_polys: u64, width: u32, points: *const u64,
                                        n_points: u64, out: *mut u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_batch_coset_extrapolate(offset_raw: u64, codeword_length: u64, codewords: *const u64, n_codewords: u64,
                                        width: u32, points: *const u64, n_points: u64, out: *mut u64) -> c_int;
    pub fn tf21_batch_coset_extrapolate_dev(offset_raw: u64, codeword_length: u64, codewords: *const u64,
                                            n_codewords: u64, width: u32, points: *const u64, n_points: u64,
                                            out: *mut u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_tip5_sample_indices(state: *mut u64, upper_bound: u32, num_indices: u64, out: *mut u32) -> c_int;
    pub fn tf21_poly_reduce_by_ntt_friendly_modulus(coeffs: *const u64, n_coeffs: u64, width: u32, shift_ntt: *const u64,
                                                    domain_length: u64, tail_length: u64, out: *mut u64,
                                                    n_out: *mut u64) -> c_int;
    pub fn tf21_poly_clean_divide(a: *const u64, n_a: u64, b: *const u64, n_b: u64, q_out: *mut u64, n_q: *mut u64) -> c_int;
    pub fn tf21_poly_square(a: *const u64, n_a: u64, width: u32, out: *mut u64) -> c_int;
    pub fn tf21_poly_square_dev(a: *const u64, n_a: u64, width: u32, out: *mut u64, s: tf21_stream_t) -> c_int;

    pub fn tf21_tip5_permute(states: *mut u64, count: u64) -> c_int;
    pub fn tf21_tip5_hash_10(input: *const u64, count: u64, out: *mut u64) -> c_int;
    pub fn tf21_tip5_hash_pairs(pairs: *const u64, count: u64, out: *mut u64) -> c_int;
    pub fn tf21_tip5_hash_varlen(input: *const u64, len: u64, out: *mut u64) -> c_int;
    pub fn tf21_tip5_hash_rows(rows: *const u64, row_len: u64, n_rows: u64, out: *mut u64) -> c_int;
    pub fn tf21_tip5_permute_dev(states: *mut u64, count: u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_tip5_hash_10_dev(input: *const u64, count: u64, out: *mut u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_tip5_hash_rows_dev(rows: *const u64, row_len: u64, n_rows: u64, out: *mut u64,
                                   s: tf21_stream_t) -> c_int;
    pub fn tf21_tip5_hash_columns_dev(cols: *const u64, n_rows: u64, n_cols: u64, col_stride_words: u64,
                                      out: *mut u64, s: tf21_stream_t) -> c_int;

    pub fn tf21_merkle_build(leafs: *const u64, n_leafs: u64, nodes_out: *mut u64) -> c_int;
    pub fn tf21_merkle_root(leafs: *const u64, n_leafs: u64, root_out: *mut u64) -> c_int;
    pub fn tf21_merkle_build_dev(leafs: *const u64, n_leafs: u64, nodes_out: *mut u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_merkle_root_dev(leafs: *const u64, n_leafs: u64, root_out: *mut u64, s: tf21_stream_t) -> c_int;
    pub fn tf21_merkle_scatter_subtree_dev(local_nodes: *const u64, n_local_leafs: u64, shard: u64, n_shards: u64,
                                           global_nodes: *mut u64, s: tf21_stream_t) -> c_int;

    pub fn tf21_ntt_sharded(data: *mut u64, n: u64, width: u32, batch: u64, inverse: c_int, n_shards: u32) -> c_int;
    pub fn tf21_merkle_build_sharded(leafs: *const u64, n_leafs: u64, nodes_out: *mut u64, n_shards: u32) -> c_int;
    pub fn tf21_merkle_auth_structure_node_indices(n_leafs: u64, leaf_indices: *const u64, n_indices: u64,
                                                   out: *mut u64, capacity: u64, count: *mut u64) -> c_int;
    pub fn tf21_merkle_authentication_structure_dev(nodes: *const u64, n_leafs: u64, leaf_indices: *const u64,
                                                    n_indices: u64, out: *mut u64, capacity: u64, count: *mut u64,
                                                    s: tf21_stream_t) -> c_int;
    pub fn tf21_merkle_authentication_structure_from_leafs(leafs: *const u64, n_leafs: u64, leaf_indices: *const u64,
                                                           n_indices: u64, out: *mut u64, capacity: u64,
                                                           count: *mut u64) -> c_int;
    pub fn tf21_mmr_peaks_from_leafs(leafs: *const u64, n_leafs: u64, peaks_out: *mut u64, n_peaks: *mut u64) -> c_int;
    pub fn tf21_mmr_peaks_from_leafs_dev(leafs: *const u64, n_leafs: u64, peaks_out: *mut u64, n_peaks: *mut u64,
                                         s: tf21_stream_t) -> c_int;
    pub fn tf21_mmr_bag_peaks(peaks: *const u64, n_peaks: u64, leaf_count: u64, out: *mut u64) -> c_int;
    pub fn tf21_mmr_bag_peaks_dev(peaks: *const u64, n_peaks: u64, leaf_count: u64, out: *mut u64,
                                  s: tf21_stream_t) -> c_int;
}
