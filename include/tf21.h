/*
 * tf21.h -- C ABI of the B200-native (sm_100a) replacement for twenty-first's STARK hot path.
 *
 * This is the drop-in boundary: a Rust shim keeps the reference's public signatures
 * (`math::ntt::{ntt,intt}`, `Polynomial::fast_coset_{evaluate,interpolate}`, `Tip5::*`,
 * `MerkleTree::{par_new,sequential_new,par_frugal_root,sequential_frugal_root}`) and calls the
 * entry points below (see INTEGRATION.md for the binding).  Reference paths are relative to
 * twenty-first/src/ of Neptune-Crypto/twenty-first v2.0.2.
 *
 * Conventions
 *  - Every word is the raw in-memory u64 of a `BFieldElement` (Montgomery form, canonical,
 *    b_field_element.rs:84-86).  `width` = 1 for `[BFieldElement]`, 3 for `[XFieldElement]`
 *    (`#[repr(transparent)] [BFieldElement; 3]`, x_field_element.rs:56-59).  A Digest is 5 words
 *    (tip5/digest.rs:28-29).
 *  - Return value: 0 on success, a negative TF21_E_* code otherwise.  Nothing aborts or throws
 *    across the boundary; the shim turns codes back into the reference's behaviour (panic for
 *    NTT length violations, Err(MerkleTreeError::..) for Merkle).
 *  - The caller owns every buffer; the library keeps no pointer past return.
 *  - Host-pointer entry points copy host->device->host internally on the calling thread's
 *    current CUDA device.  `*_dev` entry points take device pointers plus a stream
 *    (cudaStream_t passed as void*; NULL = default stream) and are asynchronous.
 *  - Alignment: device pointers are 8-byte aligned u64 arrays; the Tip5 / Merkle `*_dev` entry points
 *    (permute, hash_10, merkle_build, merkle_root, mmr_peaks_from_leafs) additionally need their
 *    INPUT arrays (and the Merkle node array) 16-byte aligned -- cudaMalloc / tf21_malloc / torch
 *    allocations are; an 8-byte-only view returns TF21_E_BAD_ARG.  The NTT / coset entry points accept
 *    any 8-byte aligned pointer.
 *  - Thread safe: tables are built once per (device, size) under a lock; scratch is per call
 *    or per (device, stream) cache guarded by a lock.
 *  - There is no CPU fallback: without a usable CUDA device every call returns TF21_E_CUDA.
 */
#ifndef TF21_H
#define TF21_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    TF21_OK = 0,
    TF21_E_LEN_NOT_POW2 = -1,    /* ntt.rs:137 `assert!(len == 0 || is_power_of_two)` (panic)        */
    TF21_E_LEN_TOO_LARGE = -2,   /* ntt.rs:135-136 len > u32::MAX (panic); also the 2^30 device limit */
    TF21_E_TOO_FEW_LEAFS = -3,   /* MerkleTreeError::TooFewLeafs, merkle_tree.rs:394-396            */
    TF21_E_INCORRECT_NUMBER_OF_LEAFS = -4, /* MerkleTreeError::IncorrectNumberOfLeafs, :398-401     */
    TF21_E_ORDER_LE_DEGREE = -5, /* polynomial.rs:1388-1392 assert (panic)                           */
    TF21_E_ALLOC = -6,           /* MerkleTreeError::TreeTooHigh, merkle_tree.rs:403-410 / cudaMalloc */
    TF21_E_CUDA = -7,            /* any CUDA runtime failure; see tf21_last_cuda_error()             */
    TF21_E_BAD_ARG = -8,         /* width not in {1,3}, NULL pointer with non-zero size, ...         */
    TF21_E_LEAF_INDEX_INVALID = -9, /* MerkleTreeError::LeafIndexInvalid, merkle_tree.rs:487-489       */
    TF21_E_CAPACITY = -10,       /* output buffer smaller than the result; *count holds the need     */
    TF21_E_DIVISION_BY_ZERO = -11, /* polynomial.rs:556-559 `expect("divisor should be non-zero")` (panic) */
    TF21_E_NCCL = -12,           /* NCCL failure in a sharded entry point; see tf21_last_cuda_error() */
};

typedef void *tf21_stream_t; /* cudaStream_t */

/* ---- runtime -------------------------------------------------------------------------------- */
/* Declares how many of the visible devices the sharded entry points use (0 = all) and prepares the
 * calling thread's current device (constant tables).  Other devices, and the NCCL communicators of the
 * sharded Merkle build (ncclCommInitAll over devices 0..n-1), are set up on first use.               */
int tf21_init(int n_devices);
/* Selects `device` for the calling thread (cudaSetDevice) and prepares it: what a one-process-per-GPU
 * launch calls with its local rank.                                                                  */
int tf21_set_device(int device);
/* Number of devices the sharded entry points will use (min(tf21_init's n_devices, visible)).         */
int tf21_device_count(void);
/* Frees cached tables and scratch of every device touched by this process.                     */
int tf21_shutdown(void);
const char *tf21_strerror(int code);
const char *tf21_last_cuda_error(void);
/* Number of kernels launched by this library in this process since start (bench bookkeeping).  */
uint64_t tf21_kernel_launch_count(void);
/* Per-launch device timing for roofline reporting: when enabled every kernel launch is bracketed
 * by CUDA events on its stream.  tf21_profile_read writes "kernel_name milliseconds\n" per
 * launch (launch order) and returns the buffer size needed; enable(0/1) also clears the log.   */
int tf21_profile_enable(int on);
int64_t tf21_profile_read(char *buf, uint64_t buflen);
/* Stream-ordered device memory helpers so a host language needs no CUDA binding of its own.   */
int tf21_malloc(void **dptr, uint64_t bytes);
int tf21_free(void *dptr);
int tf21_memcpy_h2d(void *dst_dev, const void *src_host, uint64_t bytes, tf21_stream_t stream);
int tf21_memcpy_d2h(void *dst_host, const void *src_dev, uint64_t bytes, tf21_stream_t stream);
int tf21_stream_sync(tf21_stream_t stream);

/* Diagnostic: element-wise device field arithmetic on raw words (used by the parity tests to pin
 * the Goldilocks primitives): op 0 add, 1 sub, 2 mul (all mod p, canonical result), 3 canonicalise,
 * 4 weak add then canonicalise, 5 reduce a 96-bit value a + (b & 0xffffffff) * 2^64,
 * 6 lazy add (a any, b <= p) then canonicalise, 7 canonicalise (IMAD.WIDE form), 8 lazy sub (a any,
 * b < p) then canonicalise, 9 the butterfly form of the lazy sub (b <= p), 100+S: a * 2^S mod p for the shift twiddles S (S % 3 == 0 or S in
 * {1,31,65,95}), canonical result for any a.                                                    */
int tf21_selftest_field_dev(int op, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out,
                            uint64_t n, tf21_stream_t stream);

/* Diagnostic: lands n_tiles [1024 rows][4 words] tiles of a [1024][inner_words] device matrix (16-byte aligned)
 * by TMA and returns the raw shared-memory image of each (4096 words), which pins the swizzled tile layout. */
int tf21_selftest_tma_tile_dev(const uint64_t *d_matrix, uint64_t inner_words, uint64_t n_tiles,
                               uint64_t *d_out, tf21_stream_t stream);

/* ---- NTT: math::ntt::ntt / intt (ntt.rs:67-82, 109-125) ------------------------------------- */
/* In place over `batch` contiguous arrays of n*width words. n == 0 or 1 is a no-op.
 * out[i] = sum_j x[j] * omega_n^(i j), natural order in and out; intt also multiplies by n^-1.
 * Host memory: every entry point without a `_dev` suffix takes HOST pointers of any kind.  Pinned / registered
 * memory is copied directly (both PCIe directions and the kernels overlapped); pageable memory -- a plain Vec or
 * malloc block -- goes through the library's pinned staging ring (helper threads copy between the caller's slice
 * and the ring; same results, about half the pinned rate instead of the driver's one-direction-at-a-time staging). */
int tf21_ntt(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch);
int tf21_intt(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch);
int tf21_ntt_dev(uint64_t *d_data, uint64_t n, uint32_t width, uint64_t batch, int inverse,
                 tf21_stream_t stream);

/* ---- coset evaluate / interpolate: math::polynomial (polynomial.rs:760-773,1374-1399,1907-1918) */
/* out[0..order*width) = NTT_order( zero_extend( c[i] * offset^i ) ); requires order > degree.    */
int tf21_coset_evaluate(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t width,
                        uint64_t offset_raw, uint64_t order, uint64_t *out);
/* coeffs_out[i] = iNTT_n(values)[i] * offset^-i ; all n coefficients are returned.               */
int tf21_coset_interpolate(const uint64_t *values, uint64_t n, uint32_t width, uint64_t offset_raw,
                           uint64_t *coeffs_out);
/* Fused low-degree extension = interpolate on (offset_in, n_in) then evaluate on
 * (offset_out, n_out), n_out >= n_in.                                                           */
int tf21_coset_lde(const uint64_t *values, uint64_t n_in, uint64_t offset_in_raw, uint64_t n_out,
                   uint64_t offset_out_raw, uint32_t width, uint64_t *out);
/* Device variants. `d_out` must not alias the input. The degree check of evaluate is skipped on
 * device (the caller guarantees n_coeffs <= order); n_coeffs > order returns ORDER_LE_DEGREE.   */
int tf21_coset_evaluate_dev(const uint64_t *d_coeffs, uint64_t n_coeffs, uint32_t width,
                            uint64_t offset_raw, uint64_t order, uint64_t *d_out,
                            tf21_stream_t stream);
int tf21_coset_interpolate_dev(const uint64_t *d_values, uint64_t n, uint32_t width,
                               uint64_t offset_raw, uint64_t *d_coeffs_out, tf21_stream_t stream);
int tf21_coset_lde_dev(const uint64_t *d_values, uint64_t n_in, uint64_t offset_in_raw,
                       uint64_t n_out, uint64_t offset_out_raw, uint32_t width, uint64_t *d_out,
                       tf21_stream_t stream);

/* ---- next wave: Polynomial::fast_multiply (polynomial.rs:900-932) ------------- */
/* out[0 .. n_a + n_b - 1) = a * b for two coefficient slices of the same width (1 = BFieldElement,
 * 3 = XFieldElement): zero-extended NTTs of size next_power_of_two(n_a + n_b - 1), Hadamard product
 * (Montgomery-correct), inverse NTT, truncation.  n_a == 0 or n_b == 0 is the zero polynomial:
 * nothing is written.                                                                           */
int tf21_poly_mul(const uint64_t *a, uint64_t n_a, const uint64_t *b, uint64_t n_b, uint32_t width,
                  uint64_t *out);
int tf21_poly_mul_dev(const uint64_t *d_a, uint64_t n_a, const uint64_t *d_b, uint64_t n_b, uint32_t width,
                      uint64_t *d_out, tf21_stream_t stream);

/* Polynomial::fast_square (polynomial.rs:780-802): out[0 .. 2 n_a - 1) = a * a with one forward transform
 * (tf21_poly_mul_dev does the same when both operands are the same buffer).                          */
int tf21_poly_square(const uint64_t *a, uint64_t n_a, uint32_t width, uint64_t *out);
int tf21_poly_square_dev(const uint64_t *d_a, uint64_t n_a, uint32_t width, uint64_t *d_out,
                         tf21_stream_t stream);

/* Polynomial::reduce_by_ntt_friendly_modulus (polynomial.rs:1087-1148): f mod (X^domain_length + shift(X)) with
 * deg shift < tail_length and shift given as its NTT over domain_length points (shift_factor_ntt_with_tail_length,
 * :1051-1074).  out receives min(n_coeffs, domain_length) elements; *n_out is set to that count.               */
int tf21_poly_reduce_by_ntt_friendly_modulus(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t width,
                                             const uint64_t *shift_ntt, uint64_t domain_length,
                                             uint64_t tail_length, uint64_t *out, uint64_t *n_out);
/* Polynomial<BFieldElement>::clean_divide (polynomial.rs:2358-2413): q = a / b when the division leaves no
 * remainder (the caller's promise, as in the reference; otherwise the result is unspecified).  Trailing zero
 * coefficients are ignored; *n_q = deg a - deg b + 1 coefficients are written (0 if deg a < deg b).        */
int tf21_poly_clean_divide(const uint64_t *a, uint64_t n_a, const uint64_t *b, uint64_t n_b,
                           uint64_t *q_out, uint64_t *n_q);

/* Polynomial::evaluate (polynomial.rs:309-319) of n_polys device-resident polynomials of n coefficients each
 * in n_points points (host array, same width as the coefficients): out[(poly * n_points + point) * width ..].  */
int tf21_poly_evaluate_batch_dev(const uint64_t *d_polys, uint64_t n, uint64_t n_polys, uint32_t width,
                                 const uint64_t *points, uint64_t n_points, uint64_t *d_out,
                                 tf21_stream_t stream);
/* Polynomial::{batch_,par_batch_}coset_extrapolate (polynomial.rs:2117-2331): n_codewords codewords of
 * codeword_length (a power of two) values on the coset offset * <omega>, extrapolated to n_points points;
 * out[(codeword * n_points + point) * width ..] like the reference's flat_map.                          */
int tf21_batch_coset_extrapolate(uint64_t offset_raw, uint64_t codeword_length, const uint64_t *codewords,
                                 uint64_t n_codewords, uint32_t width, const uint64_t *points,
                                 uint64_t n_points, uint64_t *out);
int tf21_batch_coset_extrapolate_dev(uint64_t offset_raw, uint64_t codeword_length,
                                     const uint64_t *d_codewords, uint64_t n_codewords, uint32_t width,
                                     const uint64_t *points, uint64_t n_points, uint64_t *d_out,
                                     tf21_stream_t stream);

/* ---- Tip5 (tip5/mod.rs:529-533, 559-586, 617-623; sponge.rs:41-56) -------------------------- */
/* Tip5::sample_indices (tip5/mod.rs:636-656): `state` = the sponge's 16 raw words, updated in place;
 * upper_bound must be a power of two (TF21_E_LEN_NOT_POW2 <-> the reference's assert).                  */
int tf21_tip5_sample_indices(uint64_t *state, uint32_t upper_bound, uint64_t num_indices, uint32_t *out);
int tf21_tip5_permute(uint64_t *states /*16 words each*/, uint64_t count);
int tf21_tip5_hash_10(const uint64_t *in /*10 words each*/, uint64_t count, uint64_t *out /*5 each*/);
int tf21_tip5_hash_pairs(const uint64_t *pairs /*left(5)|right(5)*/, uint64_t count, uint64_t *out);
/* hash_varlen of one sequence (sequential sponge; runs on the device, latency bound).          */
int tf21_tip5_hash_varlen(const uint64_t *in, uint64_t len, uint64_t out[5]);
/* hash_varlen of `n_rows` rows of `row_len` words each (row-major) -> n_rows digests: the step
 * that turns NTT codewords into Merkle leaves in the crate's callers.                          */
int tf21_tip5_hash_rows(const uint64_t *rows, uint64_t row_len, uint64_t n_rows, uint64_t *out);
int tf21_tip5_permute_dev(uint64_t *d_states, uint64_t count, tf21_stream_t stream);
int tf21_tip5_hash_10_dev(const uint64_t *d_in, uint64_t count, uint64_t *d_out, tf21_stream_t stream);
int tf21_tip5_hash_rows_dev(const uint64_t *d_rows, uint64_t row_len, uint64_t n_rows,
                            uint64_t *d_out, tf21_stream_t stream);
/* Same for column-major data (a batch of NTT codewords as tf21_ntt_dev leaves them): digest i =
 * hash_varlen(col_0[i], col_1[i], .., col_{n_cols-1}[i]) with col_c = d_cols + c * col_stride_words.
 * This is the step between the NTT codewords and MerkleTree::par_new in the crate's callers
 * (SURVEY.md 8f-1); loads are coalesced across rows.                                            */
int tf21_tip5_hash_columns_dev(const uint64_t *d_cols, uint64_t n_rows, uint64_t n_cols,
                               uint64_t col_stride_words, uint64_t *d_out, tf21_stream_t stream);

/* ---- Merkle tree (merkle_tree.rs:149-222, 299-364, 393-429) -------------------------------- */
/* nodes_out has 2*n_leafs digests: nodes[0] = 0, nodes[1] = root,
 * nodes[i] = hash_pair(nodes[2i], nodes[2i+1]), nodes[n..2n) = leafs.                          */
int tf21_merkle_build(const uint64_t *leafs, uint64_t n_leafs, uint64_t *nodes_out);
int tf21_merkle_root(const uint64_t *leafs, uint64_t n_leafs, uint64_t root_out[5]);
int tf21_merkle_build_dev(const uint64_t *d_leafs, uint64_t n_leafs, uint64_t *d_nodes_out,
                          tf21_stream_t stream);
int tf21_merkle_root_dev(const uint64_t *d_leafs, uint64_t n_leafs, uint64_t *d_root_out,
                         tf21_stream_t stream);
/* Multi-GPU assembly (subtrees shard independently, merkle_tree.rs:247-275): rank `shard` of
 * `n_shards` builds its local tree with tf21_merkle_build_dev over its n_leafs/n_shards leaves;
 * after the cap (the n_shards local roots) has been gathered, the top of the tree is
 * tf21_merkle_build_dev over those roots.  This scatters a local tree into its positions in the
 * global heap-indexed node array (only needed when one device wants the whole array).          */
int tf21_merkle_scatter_subtree_dev(const uint64_t *d_local_nodes, uint64_t n_local_leafs,
                                    uint64_t shard, uint64_t n_shards, uint64_t *d_global_nodes,
                                    tf21_stream_t stream);

/* ---- single-process sharding over the visible devices ---------------------------- */
/* `batch` columns split over n_shards host threads, shard s on device s % (visible devices); no exchange.  */
int tf21_ntt_sharded(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch, int inverse,
                     uint32_t n_shards);
/* n_shards (a power of two) subtrees built independently, then the top log2(n_shards) levels from the shard
 * roots (the tree cap); nodes_out as tf21_merkle_build.  With one shard per device (n_shards <= device count)
 * the cap is exchanged by one ncclAllGather of 40 bytes per device over NVLink and every device hashes the top
 * levels (TF21_E_NCCL on a NCCL failure); with more shards than devices, or without libnccl, the cap is
 * assembled in host memory.                                                                                */
int tf21_merkle_build_sharded(const uint64_t *leafs, uint64_t n_leafs, uint64_t *nodes_out,
                              uint32_t n_shards);
/* 1 if tf21_merkle_build_sharded would gather the cap over NCCL for this shard count, 0 for the host path. */
int tf21_sharded_uses_nccl(uint32_t n_shards);

/* ---- next wave: authentication structures (merkle_tree.rs:449-542, 614-622) ---- */
/* MerkleTree::authentication_structure_node_indices: node indices needed to prove the given leaf
 * indices, descending, de-duplicated.  Pure host logic (no device).  *count is always set to the number
 * of indices; TF21_E_CAPACITY if it exceeds `capacity` (call with out = NULL, capacity = 0 to size).   */
int tf21_merkle_auth_structure_node_indices(uint64_t n_leafs, const uint64_t *leaf_indices,
                                            uint64_t n_indices, uint64_t *out, uint64_t capacity,
                                            uint64_t *count);
/* MerkleTree::authentication_structure of a device-resident node array (as tf21_merkle_build_dev leaves
 * it): d_out[k] = nodes[index_k], 5 words each.  leaf_indices is a host array.                         */
int tf21_merkle_authentication_structure_dev(const uint64_t *d_nodes, uint64_t n_leafs,
                                             const uint64_t *leaf_indices, uint64_t n_indices,
                                             uint64_t *d_out, uint64_t capacity, uint64_t *count,
                                             tf21_stream_t stream);
/* MerkleTree::{sequential,par}_authentication_structure_from_leafs: host leaves in, host digests out. */
int tf21_merkle_authentication_structure_from_leafs(const uint64_t *leafs, uint64_t n_leafs,
                                                    const uint64_t *leaf_indices, uint64_t n_indices,
                                                    uint64_t *out, uint64_t capacity, uint64_t *count);

/* ---- next wave: MMR bulk operations (mmr/mmr_accumulator.rs:96-115, 379-391) ------- */
/* MmrAccumulator::peaks_from_leafs for any leaf count: *n_peaks = popcount(n_leafs) digests are written
 * (at most 64).  n_leafs == 0 writes nothing.                                                          */
int tf21_mmr_peaks_from_leafs(const uint64_t *leafs, uint64_t n_leafs, uint64_t *peaks_out,
                              uint64_t *n_peaks);
int tf21_mmr_peaks_from_leafs_dev(const uint64_t *d_leafs, uint64_t n_leafs, uint64_t *d_peaks_out,
                                  uint64_t *n_peaks, tf21_stream_t stream);
/* bag_peaks(peaks, leaf_count): hash_10 of the encoded leaf count folded over the peaks from the last. */
int tf21_mmr_bag_peaks(const uint64_t *peaks, uint64_t n_peaks, uint64_t leaf_count, uint64_t out[5]);
int tf21_mmr_bag_peaks_dev(const uint64_t *d_peaks, uint64_t n_peaks, uint64_t leaf_count,
                           uint64_t *d_out, tf21_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TF21_H */
