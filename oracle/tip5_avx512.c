/*
 * oracle/tip5_avx512.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-intrinsics restatement of the reference's AVX-512 IFMA/VBMI Tip5 round (twenty-first/src/tip5/avx512.rs:13-372),
 * the code path the crate itself uses on hosts with avx512f + avx512bw + avx512ifma + avx512vbmi
 * (tip5/mod.rs:36-46).  It exists so that the CPU baseline next to the GPU numbers is the reference's FASTEST
 * Tip5, not only its scalar one (oracle.c restates that: mds_generated, tip5/mod.rs:175-506).  Compiled as its own
 * translation unit with the four -mavx512* flags; oracle.c calls it only after __builtin_cpu_supports says the
 * host has them.  Pinned by the same known-answer tests as the scalar form (tests/test_oracle_kat.py) -- the
 * reference keeps those snapshots precisely to pin its two builds against each other (tip5/mod.rs:1285-1295).
 *
 * Layout notes (avx512.rs): a = state[0..8), b = state[8..16) as eight 64-bit lanes each.
 *   S-box   (:20-65):   lanes 0..3 byte-wise LOOKUP_TABLE through four masked vpermb over 64-byte slices;
 *                       lanes 4..15 x^7 = (x * x^2) * x^4 with 48/16-bit split operands and madd52 (:309-372),
 *                       reduced by reduce3x48 (:178-217).
 *   MDS+RC  (:69-175):  32-bit halves of every lane broadcast, 16-bit matrix entries, madd52lo accumulation into
 *                       accumulators seeded with the round-constant halves, one reduce2x32 (:264-306) per half.
 */
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#define P 0xffffffff00000001ULL

static uint64_t g_mds_t[16][8] __attribute__((aligned(64))); /* rows 2i, 2i+1: coefficients of input lane i / 8+i */
static uint64_t g_rc_u[80] __attribute__((aligned(64)));
static uint64_t g_rc_l[80] __attribute__((aligned(64)));
static uint8_t g_lut[256] __attribute__((aligned(64)));
static int g_ready = 0;

/* tables are derived from the same data oracle.c holds (MDS first column tip5/mod.rs:154-157, raw round
 * constants = ROUND_CONSTANTS * 2^64 mod p :68-149, LOOKUP_TABLE :50-64) instead of being transcribed a second time */
void oracle_tip5_avx512_setup(const uint64_t mds_first_column[16], const uint64_t rc_raw[80], const uint8_t lut[256]) {
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 8; k++) {
            g_mds_t[2 * i][k] = mds_first_column[(16 + k - i) % 16];         /* out k   <- in i  */
            g_mds_t[2 * i + 1][k] = mds_first_column[(16 + 8 + k - i) % 16]; /* out 8+k <- in i  (and out k <- in 8+i) */
        }
    for (int i = 0; i < 80; i++) {
        g_rc_u[i] = rc_raw[i] >> 32;
        g_rc_l[i] = rc_raw[i] & 0xffffffffULL;
    }
    memcpy(g_lut, lut, 256);
    g_ready = 1;
}

/* a + b 2^48 + c 2^96 -> F_p   (avx512.rs:178-217) */
static inline __m512i reduce3x48(__m512i a, __m512i b, __m512i c) {
    const __m512i mask32 = _mm512_set1_epi64(0xffffffffLL);
    const __m512i mask48 = _mm512_set1_epi64(0xffffffffffffLL);
    /* propagate carries */
    const __m512i ova = _mm512_srli_epi64(a, 48);
    const __m512i ovb = _mm512_srli_epi64(b, 48);
    b = _mm512_add_epi64(b, ova);
    c = _mm512_add_epi64(c, ovb);
    a = _mm512_and_epi64(a, mask48);
    b = _mm512_and_epi64(b, mask48);
    /* low 64 bits = a | b << 48;  2^96 = -1  =>  - c;  bits 64.. of b 2^48 (= b >> 16) times 2^64 = 2^32 - 1 */
    const __m512i ab = _mm512_or_epi64(a, _mm512_slli_epi64(b, 48));
    const __m512i t0 = _mm512_sub_epi64(ab, c);
    __mmask8 ov = _mm512_cmp_epu64_mask(ab, t0, _MM_CMPINT_LT);
    const __m512i bh = _mm512_srli_epi64(b, 16);
    const __m512i t2 = _mm512_sub_epi64(_mm512_slli_epi64(bh, 32), bh);
    const __m512i t1 = _mm512_mask_sub_epi64(t0, ov, t0, mask32);
    const __m512i r = _mm512_add_epi64(t1, t2);
    ov = _mm512_cmp_epu64_mask(r, t1, _MM_CMPINT_LT);
    return _mm512_mask_add_epi64(r, ov, r, mask32);
}

/* eight products x * y mod p, operands split 48 | 16 bits (avx512.rs:309-342) */
static inline __m512i mul8(__m512i x, __m512i y) {
    const __m512i mask48 = _mm512_set1_epi64(0xffffffffffffLL);
    const __m512i z = _mm512_setzero_si512();
    const __m512i xhi = _mm512_srli_epi64(x, 48), yhi = _mm512_srli_epi64(y, 48);
    const __m512i xlo = _mm512_and_epi64(x, mask48), ylo = _mm512_and_epi64(y, mask48);
    __m512i a0 = _mm512_madd52lo_epu64(z, xlo, ylo);
    __m512i b0 = _mm512_madd52lo_epu64(z, xhi, ylo);
    b0 = _mm512_madd52lo_epu64(b0, xlo, yhi);
    __m512i b4 = _mm512_madd52hi_epu64(z, xlo, ylo);
    __m512i c0 = _mm512_madd52lo_epu64(z, xhi, yhi);
    __m512i c4 = _mm512_madd52hi_epu64(z, xhi, ylo);
    c4 = _mm512_madd52hi_epu64(c4, xlo, yhi);
    b0 = _mm512_add_epi64(b0, _mm512_slli_epi64(b4, 4)); /* bit 52 of a product is bit 4 of the next 48-bit limb */
    c0 = _mm512_add_epi64(c0, _mm512_slli_epi64(c4, 4));
    return reduce3x48(a0, b0, c0);
}

/* eight squares (avx512.rs:344-372) */
static inline __m512i square8(__m512i x) {
    const __m512i mask48 = _mm512_set1_epi64(0xffffffffffffLL);
    const __m512i z = _mm512_setzero_si512();
    const __m512i xhi = _mm512_srli_epi64(x, 48), xlo = _mm512_and_epi64(x, mask48);
    const __m512i a0 = _mm512_madd52lo_epu64(z, xlo, xlo);
    __m512i b1 = _mm512_madd52lo_epu64(z, xhi, xlo);
    __m512i b4 = _mm512_madd52hi_epu64(z, xlo, xlo);
    __m512i c0 = _mm512_madd52lo_epu64(z, xhi, xhi);
    __m512i c5 = _mm512_madd52hi_epu64(z, xhi, xlo);
    const __m512i b0 = _mm512_add_epi64(_mm512_slli_epi64(b1, 1), _mm512_slli_epi64(b4, 4));
    c0 = _mm512_add_epi64(c0, _mm512_slli_epi64(c5, 5));
    return reduce3x48(a0, b0, c0);
}

/* lo + hi 2^32 -> [0, p), both limbs at most 53 bits (avx512.rs:264-306) */
static inline __m512i reduce2x32(__m512i lo, __m512i hi) {
    const __m512i u32max = _mm512_set1_epi64(0xffffffffLL);
    const __m512i x0 = _mm512_and_epi64(lo, u32max);
    const __m512i xt = _mm512_add_epi64(_mm512_srli_epi64(lo, 32), hi);
    const __m512i x1 = _mm512_and_epi64(xt, u32max);
    const __m512i x2 = _mm512_srli_epi64(xt, 32);
    const __m512i s = _mm512_add_epi64(x1, x2);
    const __m512i r = _mm512_sub_epi64(_mm512_add_epi64(_mm512_slli_epi64(s, 32), x0), x2);
    const __m512i p = _mm512_set1_epi64((long long)P);
    const __mmask8 m = _mm512_cmp_epu64_mask(r, p, _MM_CMPINT_NLT) | _mm512_cmp_epu64_mask(s, u32max, _MM_CMPINT_NLE);
    return _mm512_mask_sub_epi64(r, m, r, p);
}

static inline void sbox_layer(uint64_t *state) {
    const __m512i a = _mm512_loadu_si512(state), b = _mm512_loadu_si512(state + 8);
    /* split-and-lookup on every byte of a (only lanes 0..3 are kept): four 64-entry slices of the table */
    const __m512i c64 = _mm512_set1_epi8(0x40);
    const __m512i s0 = _mm512_loadu_si512(g_lut), s1 = _mm512_loadu_si512(g_lut + 64);
    const __m512i s2 = _mm512_loadu_si512(g_lut + 128), s3 = _mm512_loadu_si512(g_lut + 192);
    const __m512i i0 = a, i1 = _mm512_sub_epi8(i0, c64), i2 = _mm512_sub_epi8(i1, c64), i3 = _mm512_sub_epi8(i2, c64);
    __m512i sb = _mm512_setzero_si512();
    sb = _mm512_mask_permutexvar_epi8(sb, _mm512_cmplt_epu8_mask(i0, c64), i0, s0);
    sb = _mm512_mask_permutexvar_epi8(sb, _mm512_cmplt_epu8_mask(i1, c64), i1, s1);
    sb = _mm512_mask_permutexvar_epi8(sb, _mm512_cmplt_epu8_mask(i2, c64), i2, s2);
    sb = _mm512_mask_permutexvar_epi8(sb, _mm512_cmplt_epu8_mask(i3, c64), i3, s3);
    /* seventh power on all sixteen lanes (lanes 0..3 of the result are discarded) */
    const __m512i a2 = square8(a), b2 = square8(b);
    const __m512i a4 = square8(a2), b4 = square8(b2);
    const __m512i a7 = mul8(mul8(a, a2), a4), b7 = mul8(mul8(b, b2), b4);
    _mm512_storeu_si512(state, _mm512_mask_blend_epi64(0x0f, a7, sb));
    _mm512_storeu_si512(state + 8, b7);
}

static inline void mds_rcs(uint64_t *state, int round) {
    const uint32_t *a32 = (const uint32_t *)state, *b32 = (const uint32_t *)(state + 8);
    __m512i r0lo = _mm512_loadu_si512(g_rc_l + 16 * round), r1lo = _mm512_loadu_si512(g_rc_l + 16 * round + 8);
    __m512i r0hi = _mm512_loadu_si512(g_rc_u + 16 * round), r1hi = _mm512_loadu_si512(g_rc_u + 16 * round + 8);
    for (int i = 0; i < 8; i++) {
        const __m512i c0 = _mm512_load_si512(g_mds_t[2 * i]), c1 = _mm512_load_si512(g_mds_t[2 * i + 1]);
        const __m512i d0lo = _mm512_set1_epi64(a32[2 * i]), d0hi = _mm512_set1_epi64(a32[2 * i + 1]);
        const __m512i e0lo = _mm512_set1_epi64(b32[2 * i]), e0hi = _mm512_set1_epi64(b32[2 * i + 1]);
        r0lo = _mm512_madd52lo_epu64(r0lo, c0, d0lo);
        r0hi = _mm512_madd52lo_epu64(r0hi, c0, d0hi);
        r1lo = _mm512_madd52lo_epu64(r1lo, c1, d0lo);
        r1hi = _mm512_madd52lo_epu64(r1hi, c1, d0hi);
        r0lo = _mm512_madd52lo_epu64(r0lo, c1, e0lo);
        r0hi = _mm512_madd52lo_epu64(r0hi, c1, e0hi);
        r1lo = _mm512_madd52lo_epu64(r1lo, c0, e0lo);
        r1hi = _mm512_madd52lo_epu64(r1hi, c0, e0hi);
    }
    _mm512_storeu_si512(state, reduce2x32(r0lo, r0hi));
    _mm512_storeu_si512(state + 8, reduce2x32(r1lo, r1hi));
}

/* Tip5::round of the AVX-512 build (avx512.rs:13-18) */
void oracle_tip5_avx512_round(uint64_t state[16], int round) {
    sbox_layer(state);
    mds_rcs(state, round);
}

/* Tip5::permutation (tip5/mod.rs:529-533) with the AVX-512 round */
void oracle_tip5_avx512_permutation(uint64_t state[16]) {
    for (int r = 0; r < 5; r++) {
        sbox_layer(state);
        mds_rcs(state, r);
    }
}

int oracle_tip5_avx512_ready(void) { return g_ready; }
