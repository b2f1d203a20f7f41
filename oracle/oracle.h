/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatement of twenty-first's hot path; see oracle.c for the reference citations.
 * All words are raw Montgomery u64 exactly as Rust's BFieldElement stores them.
 */
#ifndef TF21_ORACLE_H
#define TF21_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_OK = 0,
    ORACLE_E_LEN_NOT_POW2 = -1,              /* ntt.rs:137 panic */
    ORACLE_E_LEN_TOO_LARGE = -2,             /* ntt.rs:135-136 panic */
    ORACLE_E_TOO_FEW_LEAFS = -3,             /* MerkleTreeError::TooFewLeafs */
    ORACLE_E_INCORRECT_NUMBER_OF_LEAFS = -4, /* MerkleTreeError::IncorrectNumberOfLeafs */
    ORACLE_E_ORDER_LE_DEGREE = -5,           /* polynomial.rs:1388-1392 panic */
    ORACLE_E_LEAF_INDEX_INVALID = -9,        /* MerkleTreeError::LeafIndexInvalid */
};

int oracle_ntt(uint64_t *x, uint64_t n, uint32_t w);
int oracle_intt(uint64_t *x, uint64_t n, uint32_t w);
int oracle_ntt_batch(uint64_t *x, uint64_t n, uint32_t w, uint64_t batch, int inverse, int threads);

void oracle_poly_scale(uint64_t *coeffs, uint64_t n_coeffs, uint32_t w, uint64_t alpha_raw);
int oracle_coset_evaluate(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t w, uint64_t offset_raw,
                          uint64_t order, uint64_t *out);
int oracle_coset_interpolate(const uint64_t *values, uint64_t n, uint32_t w, uint64_t offset_raw,
                             uint64_t *coeffs_out);
void oracle_poly_naive_multiply(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint32_t w,
                                uint64_t *out);
int oracle_poly_fast_multiply(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint32_t w,
                              uint64_t *out);
uint64_t oracle_poly_evaluate(const uint64_t *coeffs, uint64_t n_coeffs, uint64_t x_raw);
int64_t oracle_poly_reduce_by_ntt_friendly_modulus(const uint64_t *coeffs, uint64_t n, uint32_t w,
                                                   const uint64_t *shift_ntt, uint64_t domain_length,
                                                   uint64_t tail_length, uint64_t *out);
int64_t oracle_poly_naive_divide(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint64_t *quot,
                                 uint64_t *rem);
void oracle_poly_evaluate_w(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t w, const uint64_t *x, uint64_t *out);
int oracle_batch_coset_extrapolate(uint64_t offset_raw, uint64_t n, const uint64_t *codewords, uint64_t n_codewords,
                                   uint32_t w, const uint64_t *points, uint64_t n_points, uint64_t *out);
void oracle_tip5_sample_indices(uint64_t state[16], uint32_t upper_bound, uint64_t num_indices, uint32_t *out);

void oracle_tip5_permutation(uint64_t state[16]);
/* one round in the scalar build's form (mds_generated, tip5/mod.rs:175-253) and in NaiveTip5's (tip5/naive.rs:26-76) */
void oracle_tip5_hash_rows_batch(const uint64_t *rows, uint64_t row_len, uint64_t n_rows, uint64_t *out, int threads);
void oracle_tip5_round(uint64_t state[16], int round);
int oracle_tip5_avx512_available(void);
int oracle_tip5_set_impl(int impl);
int oracle_tip5_round_avx512(uint64_t state[16], int round);
void oracle_tip5_round_naive(uint64_t state[16], int round);
void oracle_tip5_hash_10(const uint64_t in[10], uint64_t out[5]);
void oracle_tip5_hash_pair(const uint64_t left[5], const uint64_t right[5], uint64_t out[5]);
void oracle_tip5_hash_varlen(const uint64_t *in, uint64_t len, uint64_t out[5]);
uint64_t oracle_tip5_hasher_bytes(const uint8_t *bytes, uint64_t n_bytes);
void oracle_tip5_permute_batch(uint64_t *states, uint64_t count, int threads);
void oracle_tip5_hash_pairs_batch(const uint64_t *pairs, uint64_t count, uint64_t *out, int threads);
void oracle_digest_to_hex(const uint64_t digest_raw[5], char out[81]);

int oracle_merkle_sequential_new(const uint64_t *leafs, uint64_t n, uint64_t *nodes);
int oracle_merkle_par_new(const uint64_t *leafs, uint64_t n, uint64_t *nodes, int num_threads,
                          uint64_t cutoff);
int oracle_merkle_sequential_frugal_root(const uint64_t *leafs, uint64_t n, uint64_t root[5]);
int oracle_merkle_par_frugal_root(const uint64_t *leafs, uint64_t n, uint64_t root[5], int num_threads,
                                  uint64_t cutoff);

uint64_t oracle_mmr_peaks_from_leafs(const uint64_t *leafs, uint64_t n, uint64_t *peaks /* <= 64 digests */);
void oracle_mmr_bag_peaks(const uint64_t *peaks, uint64_t n_peaks, uint64_t leaf_count, uint64_t out[5]);
int64_t oracle_auth_structure_node_indices(uint64_t num_leafs, const uint64_t *leaf_indices, uint64_t n_indices,
                                           uint64_t *out);

uint64_t oracle_bfe_new(uint64_t v);
uint64_t oracle_bfe_value(uint64_t raw);
uint64_t oracle_bfe_add(uint64_t a, uint64_t b);
uint64_t oracle_bfe_sub(uint64_t a, uint64_t b);
uint64_t oracle_bfe_mul(uint64_t a, uint64_t b);
uint64_t oracle_bfe_mod_pow(uint64_t a, uint64_t e);
uint64_t oracle_bfe_inverse_or_zero(uint64_t a);
uint64_t oracle_bfe_primitive_root_of_unity(uint64_t n);
void oracle_bfe_new_array(uint64_t *x, uint64_t n);
void oracle_bfe_value_array(uint64_t *x, uint64_t n);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
