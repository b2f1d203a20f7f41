/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference algorithms on the hot path
 * (Neptune-Crypto/twenty-first v2.0.2).  It is the parity checker for the CUDA path and the
 * timed CPU baseline; the product never links or calls it.  Each function cites the
 * reference file:line it follows (paths relative to twenty-first/src/).
 *
 * The reference itself (Rust, cargo) cannot be built in this image: there is no Rust
 * toolchain and no network, so there is no oracle/_ref.  Parity is pinned instead against the
 * reference's known-answer tests (see tests/test_oracle_kat.py and tests/golden/).
 *
 * All arrays hold raw Montgomery words exactly as a Rust `&[BFieldElement]` /
 * `&[XFieldElement]` / `&[Digest]` does in memory.
 */
#include "oracle.h"
#include "tip5_mds_generated.h"

#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "field.h"

/* ------------------------------------------------------------------------------------------
 * NTT  (math/ntt.rs)
 * ---------------------------------------------------------------------------------------- */

#define NUM_DOMAINS 32 /* ntt.rs:19-32 */

typedef struct {
    uint64_t **stages; /* stages[i] has 2^i entries */
    unsigned num_stages;
} twiddle_table;

static twiddle_table *g_fwd[NUM_DOMAINS + 1];
static twiddle_table *g_inv[NUM_DOMAINS + 1];
static uint32_t *g_swap[NUM_DOMAINS + 1];

/* twiddle_factors, ntt.rs:309-324 */
static twiddle_table *make_twiddles(uint32_t slice_len, uint64_t root_raw) {
    twiddle_table *t = (twiddle_table *)calloc(1, sizeof(*t));
    unsigned log2n = slice_len ? 31 - __builtin_clz(slice_len) : 0;
    t->num_stages = log2n;
    t->stages = (uint64_t **)calloc(log2n ? log2n : 1, sizeof(uint64_t *));
    for (unsigned i = 0; i < log2n; i++) {
        uint32_t m = 1u << i;
        uint32_t exponent = slice_len / (2 * m);
        uint64_t w_m = bfe_mod_pow(root_raw, exponent);
        uint64_t *w = (uint64_t *)malloc(sizeof(uint64_t) * m);
        w[0] = bfe_new(1);
        for (uint32_t j = 1; j < m; j++) w[j] = bfe_mul(w[j - 1], w_m);
        t->stages[i] = w;
    }
    return t;
}

/* bitreverse, ntt.rs:241-248 */
static inline uint32_t bitreverse(uint32_t k, uint32_t log2_n) {
    k = ((k & 0x55555555u) << 1) | ((k & 0xaaaaaaaau) >> 1);
    k = ((k & 0x33333333u) << 2) | ((k & 0xccccccccu) >> 2);
    k = ((k & 0x0f0f0f0fu) << 4) | ((k & 0xf0f0f0f0u) >> 4);
    k = ((k & 0x00ff00ffu) << 8) | ((k & 0xff00ff00u) >> 8);
    k = (k >> 16) | (k << 16);
    return k >> ((32 - log2_n) & 0x1f);
}

/* swap_indices, ntt.rs:239-284: entry k holds rev(k) iff k < rev(k), else 0 ("None") */
static uint32_t *make_swaps(uint32_t len) {
    unsigned log2n = len ? 31 - __builtin_clz(len) : 0;
    uint32_t *s = (uint32_t *)malloc(sizeof(uint32_t) * (len ? len : 1));
    for (uint32_t k = 0; k < len; k++) {
        uint32_t r = bitreverse(k, log2n);
        s[k] = (k < r) ? r : 0;
    }
    return s;
}

/* The reference caches per-size tables behind OnceLock (ntt.rs:71,113,166). */
static void ensure_tables(unsigned log2n, int inverse) {
#pragma omp critical(tf21_oracle_tables)
    {
        uint32_t n = 1u << log2n;
        if (!g_swap[log2n]) g_swap[log2n] = make_swaps(n);
        if (!inverse && !g_fwd[log2n]) {
            uint64_t omega = bfe_primitive_root_of_unity_log2(log2n);
            g_fwd[log2n] = make_twiddles(n, omega); /* ntt.rs:74-79 */
        }
        if (inverse && !g_inv[log2n]) {
            uint64_t omega = bfe_primitive_root_of_unity_log2(log2n);
            g_inv[log2n] = make_twiddles(n, bfe_inverse_or_zero(omega)); /* ntt.rs:116-121 */
        }
    }
}

/* ntt_unchecked, ntt.rs:153-215, for an element made of `w` BFieldElements on which the
 * twiddle multiply acts coefficient-wise (x_field_element.rs:540-548, 620-625) and add/sub act
 * coefficient-wise (x_field_element.rs:479-489, 560-577). */
static void ntt_unchecked(uint64_t *x, uint32_t n, uint32_t w, const twiddle_table *tw,
                          const uint32_t *swaps) {
    if (n == 0) return;
    for (uint32_t k = 0; k < n; k++) { /* ntt.rs:189-193 */
        uint32_t r = swaps[k];
        if (r) {
            for (uint32_t c = 0; c < w; c++) {
                uint64_t t = x[(uint64_t)k * w + c];
                x[(uint64_t)k * w + c] = x[(uint64_t)r * w + c];
                x[(uint64_t)r * w + c] = t;
            }
        }
    }
    uint32_t m = 1;
    for (unsigned s = 0; s < tw->num_stages; s++) { /* ntt.rs:195-214 */
        const uint64_t *twiddles = tw->stages[s];
        uint32_t k = 0;
        while (k < n) {
            for (uint32_t j = 0; j < m; j++) {
                uint64_t idx1 = (uint64_t)(k + j) * w;
                uint64_t idx2 = (uint64_t)(k + j + m) * w;
                for (uint32_t c = 0; c < w; c++) {
                    uint64_t u = x[idx1 + c];
                    uint64_t v = bfe_mul(x[idx2 + c], twiddles[j]);
                    x[idx1 + c] = bfe_add(u, v);
                    x[idx2 + c] = bfe_sub(u, v);
                }
            }
            k += 2 * m;
        }
        m *= 2;
    }
}

static int check_len(uint64_t n) {
    if (n > 0xffffffffULL) return ORACLE_E_LEN_TOO_LARGE; /* ntt.rs:135-136 (panic) */
    if (n != 0 && (n & (n - 1)) != 0) return ORACLE_E_LEN_NOT_POW2; /* ntt.rs:137 (panic) */
    return 0;
}

/* ntt, ntt.rs:67-82 */
int oracle_ntt(uint64_t *x, uint64_t n, uint32_t w) {
    int e = check_len(n);
    if (e) return e;
    if (n == 0) return 0;
    unsigned log2n = 63 - __builtin_clzll(n);
    if (log2n >= NUM_DOMAINS) return ORACLE_E_LEN_TOO_LARGE;
    ensure_tables(log2n, 0);
    ntt_unchecked(x, (uint32_t)n, w, g_fwd[log2n], g_swap[log2n]);
    return 0;
}

/* intt + unscale, ntt.rs:109-125, 220-228 */
int oracle_intt(uint64_t *x, uint64_t n, uint32_t w) {
    int e = check_len(n);
    if (e) return e;
    if (n == 0) return 0;
    unsigned log2n = 63 - __builtin_clzll(n);
    if (log2n >= NUM_DOMAINS) return ORACLE_E_LEN_TOO_LARGE;
    ensure_tables(log2n, 1);
    ntt_unchecked(x, (uint32_t)n, w, g_inv[log2n], g_swap[log2n]);
    uint64_t n_inv = bfe_inverse_or_zero(bfe_new(n)); /* ntt.rs:225 */
    for (uint64_t i = 0; i < n * w; i++) x[i] = bfe_mul(x[i], n_inv);
    return 0;
}

/* Batched form: the crate expects callers to parallelise over columns with rayon
 * (ntt.rs:250-269, benches/tip5.rs:43-49); OpenMP plays rayon's role here. */
int oracle_ntt_batch(uint64_t *x, uint64_t n, uint32_t w, uint64_t batch, int inverse,
                     int threads) {
    int e = check_len(n);
    if (e) return e;
    if (n == 0 || batch == 0) return 0;
    unsigned log2n = 63 - __builtin_clzll(n);
    if (log2n >= NUM_DOMAINS) return ORACLE_E_LEN_TOO_LARGE;
    ensure_tables(log2n, inverse);
    int rc = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (uint64_t b = 0; b < batch; b++) {
        int r = inverse ? oracle_intt(x + b * n * w, n, w) : oracle_ntt(x + b * n * w, n, w);
        if (r) rc = r;
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Polynomial coset evaluate / interpolate  (math/polynomial.rs)
 * ---------------------------------------------------------------------------------------- */

/* Polynomial::scale, polynomial.rs:760-773 (alpha is a BFieldElement; FF * BFE acts
 * coefficient-wise for XFE, x_field_element.rs:540-548) */
void oracle_poly_scale(uint64_t *coeffs, uint64_t n_coeffs, uint32_t w, uint64_t alpha_raw) {
    uint64_t power = bfe_new(1);
    for (uint64_t i = 0; i < n_coeffs; i++) {
        for (uint32_t c = 0; c < w; c++) coeffs[i * w + c] = bfe_mul(coeffs[i * w + c], power);
        power = bfe_mul(power, alpha_raw);
    }
}

/* Polynomial::fast_coset_evaluate, polynomial.rs:1374-1399.
 * `out` has order*w words.  The reference asserts order > degree (panic). */
int oracle_coset_evaluate(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t w,
                          uint64_t offset_raw, uint64_t order, uint64_t *out) {
    /* degree = index of the last non-zero coefficient, -1 for the zero polynomial
     * (polynomial.rs `degree`); Polynomial::new strips trailing zeros. */
    int64_t degree = -1;
    for (uint64_t i = n_coeffs; i-- > 0;) {
        int nz = 0;
        for (uint32_t c = 0; c < w; c++) nz |= coeffs[i * w + c] != 0;
        if (nz) {
            degree = (int64_t)i;
            break;
        }
    }
    if (!((int64_t)order > degree)) return ORACLE_E_ORDER_LE_DEGREE; /* polynomial.rs:1388-1392 */
    uint64_t live = (uint64_t)(degree + 1);
    memset(out, 0, sizeof(uint64_t) * order * w);          /* resize(order, ZERO) :1395 */
    memcpy(out, coeffs, sizeof(uint64_t) * live * w);
    oracle_poly_scale(out, live, w, offset_raw);            /* :1394 */
    return oracle_ntt(out, order, w);                       /* :1396 */
}

/* Polynomial::fast_coset_interpolate, polynomial.rs:1907-1918.  Returns all n coefficients
 * (Polynomial equality ignores trailing zeros). */
int oracle_coset_interpolate(const uint64_t *values, uint64_t n, uint32_t w, uint64_t offset_raw,
                             uint64_t *coeffs_out) {
    memcpy(coeffs_out, values, sizeof(uint64_t) * n * w); /* values.to_vec() :1912 */
    int e = oracle_intt(coeffs_out, n, w);                /* :1914 */
    if (e) return e;
    oracle_poly_scale(coeffs_out, n, w, bfe_inverse_or_zero(offset_raw)); /* :1917 */
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Tip5  (tip5/mod.rs)
 * ---------------------------------------------------------------------------------------- */

#define STATE_SIZE 16           /* tip5/mod.rs:27 */
#define NUM_SPLIT_AND_LOOKUP 4  /* :28 */
#define CAPACITY 6              /* :30 */
#define RATE 10                 /* :31 */
#define NUM_ROUNDS 5            /* :32 */
#define DIGEST_LEN 5            /* tip5/digest.rs:49 */

static uint8_t g_lut[256];
static uint64_t g_rc_raw[NUM_ROUNDS * STATE_SIZE];
static uint64_t g_mds_raw[STATE_SIZE];
static int g_tip5_ready = 0;

/* ROUND_CONSTANTS, tip5/mod.rs:68-149 (canonical values, data) */
static const uint64_t ROUND_CONSTANTS[NUM_ROUNDS * STATE_SIZE] = {
    13630775303355457758ULL, 16896927574093233874ULL, 10379449653650130495ULL,
    1965408364413093495ULL,  15232538947090185111ULL, 15892634398091747074ULL,
    3989134140024871768ULL,  2851411912127730865ULL,  8709136439293758776ULL,
    3694858669662939734ULL,  12692440244315327141ULL, 10722316166358076749ULL,
    12745429320441639448ULL, 17932424223723990421ULL, 7558102534867937463ULL,
    15551047435855531404ULL, 17532528648579384106ULL, 5216785850422679555ULL,
    15418071332095031847ULL, 11921929762955146258ULL, 9738718993677019874ULL,
    3464580399432997147ULL,  13408434769117164050ULL, 264428218649616431ULL,
    4436247869008081381ULL,  4063129435850804221ULL,  2865073155741120117ULL,
    5749834437609765994ULL,  6804196764189408435ULL,  17060469201292988508ULL,
    9475383556737206708ULL,  12876344085611465020ULL, 13835756199368269249ULL,
    1648753455944344172ULL,  9836124473569258483ULL,  12867641597107932229ULL,
    11254152636692960595ULL, 16550832737139861108ULL, 11861573970480733262ULL,
    1256660473588673495ULL,  13879506000676455136ULL, 10564103842682358721ULL,
    16142842524796397521ULL, 3287098591948630584ULL,  685911471061284805ULL,
    5285298776918878023ULL,  18310953571768047354ULL, 3142266350630002035ULL,
    549990724933663297ULL,   4901984846118077401ULL,  11458643033696775769ULL,
    8706785264119212710ULL,  12521758138015724072ULL, 11877914062416978196ULL,
    11333318251134523752ULL, 3933899631278608623ULL,  16635128972021157924ULL,
    10291337173108950450ULL, 4142107155024199350ULL,  16973934533787743537ULL,
    11068111539125175221ULL, 17546769694830203606ULL, 5315217744825068993ULL,
    4609594252909613081ULL,  3350107164315270407ULL,  17715942834299349177ULL,
    9600609149219873996ULL,  12894357635820003949ULL, 4597649658040514631ULL,
    7735563950920491847ULL,  1663379455870887181ULL,  13889298103638829706ULL,
    7375530351220884434ULL,  3502022433285269151ULL,  9231805330431056952ULL,
    9252272755288523725ULL,  10014268662326746219ULL, 15565031632950843234ULL,
    1209725273521819323ULL,  6024642864597845108ULL,
};

/* MDS_MATRIX_FIRST_COLUMN, tip5/mod.rs:154-157 */
static const uint64_t MDS_FIRST_COLUMN[STATE_SIZE] = {
    61402, 1108, 28750, 33823, 7454, 43244, 53865, 12034,
    56951, 27521, 41351, 40901, 12021, 59689, 26798, 17845,
};

/* AVX-512 IFMA/VBMI round (tip5_avx512.c <- tip5/avx512.rs), present when the build had the flags */
#ifdef TF21_ORACLE_AVX512
void oracle_tip5_avx512_setup(const uint64_t mds_first_column[16], const uint64_t rc_raw[80], const uint8_t lut[256]);
void oracle_tip5_avx512_permutation(uint64_t state[16]);
void oracle_tip5_avx512_round(uint64_t state[16], int round);
#endif
static int g_tip5_impl = 0; /* 0 = scalar (mds_generated), 1 = AVX-512 */

static void tip5_setup(void) {
    if (g_tip5_ready) return;
#pragma omp critical(tf21_oracle_tip5)
    {
        if (!g_tip5_ready) {
            /* LOOKUP_TABLE, tip5/mod.rs:50-64; regenerated from its defining formula
             * ((x+1)^3 + 256) mod 257 (test lookup_table_is_correct, tip5/mod.rs:1034-1053;
             * offset_fermat_cube_map :1022-1026) instead of being transcribed. */
            for (unsigned i = 0; i < 256; i++) {
                uint32_t xx = (i + 1) * (i + 1) * (i + 1);
                g_lut[i] = (uint8_t)((xx + 256) % 257);
            }
            for (int i = 0; i < NUM_ROUNDS * STATE_SIZE; i++) g_rc_raw[i] = bfe_new(ROUND_CONSTANTS[i]);
            for (int i = 0; i < STATE_SIZE; i++) g_mds_raw[i] = bfe_new(MDS_FIRST_COLUMN[i]);
#ifdef TF21_ORACLE_AVX512
            oracle_tip5_avx512_setup(MDS_FIRST_COLUMN, g_rc_raw, g_lut);
#endif
            g_tip5_ready = 1;
        }
    }
}

/* split_and_lookup, tip5/mod.rs:197-207: byte-wise LUT on the raw (Montgomery) LE bytes */
static inline uint64_t split_and_lookup(uint64_t raw) {
    uint64_t out = 0;
    for (int i = 0; i < 8; i++) out |= (uint64_t)g_lut[(raw >> (8 * i)) & 0xff] << (8 * i);
    return out;
}

/* sbox_layer, tip5/mod.rs:184-194 */
static inline void tip5_sbox_layer(uint64_t s[STATE_SIZE]) {
    for (int i = 0; i < NUM_SPLIT_AND_LOOKUP; i++) s[i] = split_and_lookup(s[i]);
    for (int i = NUM_SPLIT_AND_LOOKUP; i < STATE_SIZE; i++) {
        uint64_t sq = bfe_mul(s[i], s[i]);
        uint64_t qu = bfe_mul(sq, sq);
        s[i] = bfe_mul(s[i], bfe_mul(sq, qu));
    }
}

/* mds_generated, tip5/mod.rs:210-253: the raw words split into 32-bit halves, each half through the generated
 * 16-point cyclic convolution (tip5_mds_generated.h <- :256-506, outputs are 16 x the sums), recombined as
 * (lo >> 4) + (hi << 28) and folded with 2^64 = 2^32 - 1.  Like the reference it may leave a degenerate
 * representation (>= p), which the round-constant addition corrects (:222-242, tests :1122-1142). */
static inline void tip5_mds_generated(uint64_t s[STATE_SIZE]) {
    uint64_t lo[STATE_SIZE], hi[STATE_SIZE], lo2[STATE_SIZE], hi2[STATE_SIZE];
    for (int i = 0; i < STATE_SIZE; i++) {
        hi[i] = s[i] >> 32;
        lo[i] = s[i] & 0xffffffffULL;
    }
    tip5_generated_function(lo, lo2);
    tip5_generated_function(hi, hi2);
    for (int r = 0; r < STATE_SIZE; r++) {
        u128 v = (u128)(lo2[r] >> 4) + ((u128)hi2[r] << 28);
        uint64_t s_hi = (uint64_t)(v >> 64), s_lo = (uint64_t)v;
        uint64_t add = s_hi * 0xffffffffULL;
        uint64_t res = s_lo + add;
        int over = res < s_lo;
        s[r] = over ? res + 0xffffffffULL : res;
    }
}

/* the readable MDS of tip5/naive.rs:54-68 (the reference's differential oracle for mds_generated,
 * tip5/naive.rs:94-106); kept as the cross-check of the generated form */
static inline void tip5_mds_naive(uint64_t s[STATE_SIZE]) {
    uint64_t t[STATE_SIZE];
    for (int row = 0; row < STATE_SIZE; row++) {
        uint64_t acc = 0;
        for (int col = 0; col < STATE_SIZE; col++) {
            int idx = (STATE_SIZE + row - col) % STATE_SIZE;
            acc = bfe_add(acc, bfe_mul(g_mds_raw[idx], s[col]));
        }
        t[row] = acc;
    }
    for (int i = 0; i < STATE_SIZE; i++) s[i] = t[i];
}

/* One round, tip5/mod.rs:175-181: sbox_layer, mds_generated, += ROUND_CONSTANTS */
static inline void tip5_round(uint64_t s[STATE_SIZE], int round) {
    tip5_sbox_layer(s);
    tip5_mds_generated(s);
    for (int i = 0; i < STATE_SIZE; i++) s[i] = bfe_add(s[i], g_rc_raw[round * STATE_SIZE + i]);
}

/* NaiveTip5::round, tip5/naive.rs:26-76 */
void oracle_tip5_round_naive(uint64_t s[16], int round) {
    tip5_setup();
    tip5_sbox_layer(s);
    tip5_mds_naive(s);
    for (int i = 0; i < STATE_SIZE; i++) s[i] = bfe_add(s[i], g_rc_raw[round * STATE_SIZE + i]);
}

/* Tip5::round (scalar build), tip5/mod.rs:175-181 */
void oracle_tip5_round(uint64_t s[16], int round) {
    tip5_setup();
    tip5_round(s, round);
}

/* 1 if this build holds the AVX-512 round AND the host has avx512f/bw/ifma/vbmi (the reference's cfg, tip5/mod.rs:36-46) */
int oracle_tip5_avx512_available(void) {
#ifdef TF21_ORACLE_AVX512
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
           __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("avx512vbmi");
#else
    return 0;
#endif
}

/* selects the round every hash below uses: 0 = scalar build (mds_generated), 1 = AVX-512 build; returns the choice made */
int oracle_tip5_set_impl(int impl) {
    tip5_setup();
    g_tip5_impl = (impl == 1 && oracle_tip5_avx512_available()) ? 1 : 0;
    return g_tip5_impl;
}

/* Tip5::round of the AVX-512 build (tip5/avx512.rs:13-18); -1 when not available */
int oracle_tip5_round_avx512(uint64_t s[16], int round) {
#ifdef TF21_ORACLE_AVX512
    if (!oracle_tip5_avx512_available()) return -1;
    tip5_setup();
    oracle_tip5_avx512_round(s, round);
    return 0;
#else
    (void)s;
    (void)round;
    return -1;
#endif
}

/* Tip5::permutation, tip5/mod.rs:529-533 */
void oracle_tip5_permutation(uint64_t state[16]) {
    tip5_setup();
#ifdef TF21_ORACLE_AVX512
    if (g_tip5_impl == 1) {
        oracle_tip5_avx512_permutation(state);
        return;
    }
#endif
    for (int r = 0; r < NUM_ROUNDS; r++) tip5_round(state, r);
}

/* Tip5::hash_10, tip5/mod.rs:559-569 ; Tip5::new(Domain::FixedLength) :511-526 */
void oracle_tip5_hash_10(const uint64_t in[10], uint64_t out[5]) {
    uint64_t s[STATE_SIZE];
    uint64_t one = bfe_new(1);
    for (int i = 0; i < RATE; i++) s[i] = in[i];
    for (int i = RATE; i < STATE_SIZE; i++) s[i] = one;
    oracle_tip5_permutation(s);
    for (int i = 0; i < DIGEST_LEN; i++) out[i] = s[i];
}

/* Tip5::hash_pair, tip5/mod.rs:577-586 */
void oracle_tip5_hash_pair(const uint64_t left[5], const uint64_t right[5], uint64_t out[5]) {
    uint64_t in[RATE];
    for (int i = 0; i < DIGEST_LEN; i++) {
        in[i] = left[i];
        in[DIGEST_LEN + i] = right[i];
    }
    oracle_tip5_hash_10(in, out);
}

/* Tip5::hash_varlen, tip5/mod.rs:617-623 ; Sponge::pad_and_absorb_all, sponge.rs:41-56 ;
 * absorb (overwrite mode), tip5/mod.rs:684-691 */
void oracle_tip5_hash_varlen(const uint64_t *in, uint64_t len, uint64_t out[5]) {
    uint64_t s[STATE_SIZE];
    memset(s, 0, sizeof(s)); /* Domain::VariableLength */
    uint64_t full = len / RATE;
    for (uint64_t c = 0; c < full; c++) {
        for (int i = 0; i < RATE; i++) s[i] = in[c * RATE + i];
        oracle_tip5_permutation(s);
    }
    uint64_t rem = len - full * RATE;
    uint64_t last[RATE];
    memset(last, 0, sizeof(last));
    for (uint64_t i = 0; i < rem; i++) last[i] = in[full * RATE + i];
    last[rem] = bfe_new(1);
    for (int i = 0; i < RATE; i++) s[i] = last[i];
    oracle_tip5_permutation(s);
    for (int i = 0; i < DIGEST_LEN; i++) out[i] = s[i];
}

/* impl Hasher for Tip5, tip5/mod.rs:701-726: write() absorbs 8-byte LE chunks mapped through
 * BFieldElement::new, zero padded to RATE; finish() = state[0].value() */
uint64_t oracle_tip5_hasher_bytes(const uint8_t *bytes, uint64_t n_bytes) {
    uint64_t s[STATE_SIZE];
    memset(s, 0, sizeof(s)); /* Tip5::init() */
    uint64_t n_elems = (n_bytes + 7) / 8;
    for (uint64_t c = 0; c * RATE < n_elems; c++) {
        uint64_t buf[RATE];
        memset(buf, 0, sizeof(buf));
        for (uint64_t i = 0; i < RATE && c * RATE + i < n_elems; i++) {
            uint64_t v = 0;
            uint64_t off = (c * RATE + i) * 8;
            for (uint64_t b = 0; b < 8 && off + b < n_bytes; b++) v |= (uint64_t)bytes[off + b] << (8 * b);
            buf[i] = bfe_new(v);
        }
        for (int i = 0; i < RATE; i++) s[i] = buf[i];
        oracle_tip5_permutation(s);
    }
    return bfe_value(s[0]);
}

/* Batched helpers mirroring the caller-side pattern benches/tip5.rs:43-49 */
void oracle_tip5_permute_batch(uint64_t *states, uint64_t count, int threads) {
    tip5_setup();
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(static) num_threads(threads)
    for (uint64_t i = 0; i < count; i++) oracle_tip5_permutation(states + 16 * i);
}

void oracle_tip5_hash_pairs_batch(const uint64_t *pairs, uint64_t count, uint64_t *out, int threads) {
    tip5_setup();
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(static) num_threads(threads)
    for (uint64_t i = 0; i < count; i++) oracle_tip5_hash_10(pairs + 10 * i, out + 5 * i);
}

/* hash_varlen of every row of a row-major matrix (the caller-side par_iter pattern, benches/tip5.rs:43-49) */
void oracle_tip5_hash_rows_batch(const uint64_t *rows, uint64_t row_len, uint64_t n_rows, uint64_t *out, int threads) {
    tip5_setup();
#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
#else
    (void)threads;
#endif
    for (uint64_t i = 0; i < n_rows; i++) oracle_tip5_hash_varlen(rows + i * row_len, row_len, out + 5 * i);
}

/* Digest -> hex, tip5/digest.rs:85-90,144-152 + b_field_element.rs:615-621:
 * canonical value of each element, little-endian bytes, lower-case hex. out: 81 bytes. */
void oracle_digest_to_hex(const uint64_t digest_raw[5], char out[81]) {
    static const char *hexd = "0123456789abcdef";
    for (int i = 0; i < DIGEST_LEN; i++) {
        uint64_t v = bfe_value(digest_raw[i]);
        for (int b = 0; b < 8; b++) {
            uint8_t byte = (uint8_t)(v >> (8 * b));
            out[(i * 8 + b) * 2] = hexd[byte >> 4];
            out[(i * 8 + b) * 2 + 1] = hexd[byte & 15];
        }
    }
    out[80] = 0;
}

/* ------------------------------------------------------------------------------------------
 * Merkle tree  (util_types/merkle_tree.rs)
 * ---------------------------------------------------------------------------------------- */

/* initialize_merkle_tree_nodes, merkle_tree.rs:393-429 */
static int merkle_init(const uint64_t *leafs, uint64_t n, uint64_t *nodes) {
    if (n == 0) return ORACLE_E_TOO_FEW_LEAFS;                    /* :394-396 */
    if (n & (n - 1)) return ORACLE_E_INCORRECT_NUMBER_OF_LEAFS;   /* :398-401 */
    memset(nodes, 0, sizeof(uint64_t) * 5 * n);                   /* ALL_ZERO fill :415-419 */
    memcpy(nodes + 5 * n, leafs, sizeof(uint64_t) * 5 * n);       /* :426 */
    return 0;
}

/* MerkleTree::sequential_new, merkle_tree.rs:149-153 ; sequentially_fill_tree :216-222 */
int oracle_merkle_sequential_new(const uint64_t *leafs, uint64_t n, uint64_t *nodes) {
    int e = merkle_init(leafs, n, nodes);
    if (e) return e;
    tip5_setup();
    for (uint64_t i = n - 1; i >= 1; i--)
        oracle_tip5_hash_pair(nodes + 5 * (2 * i), nodes + 5 * (2 * i + 1), nodes + 5 * i);
    return 0;
}

/* MerkleTree::par_new, merkle_tree.rs:165-212.  `num_threads` plays rayon's thread count
 * (rounded down to a power of two, :376-388); `cutoff` is
 * config::merkle_tree_parallelization_cutoff (config.rs:37, default 512). Subtrees
 * (subtrees_mut :247-275) are filled bottom-up by independent workers. */
int oracle_merkle_par_new(const uint64_t *leafs, uint64_t n, uint64_t *nodes, int num_threads,
                          uint64_t cutoff) {
    int e = merkle_init(leafs, n, nodes);
    if (e) return e;
    tip5_setup();
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
    num_threads = 1;
#endif
    uint64_t threads = 1;
    while (threads * 2 <= (uint64_t)num_threads) threads *= 2; /* previous power of two :379-384 */
    if (cutoff < 2) cutoff = 2;
    uint64_t remaining = n;
    while (remaining >= cutoff) {
        while (threads > remaining / 2) threads /= 2; /* :183-185 */
        /* subtree t owns, in layer with `width` nodes, indices [width + t*width/threads, ...) */
        uint64_t sub_leaves = remaining / threads;
#pragma omp parallel for schedule(static) num_threads((int)threads)
        for (uint64_t t = 0; t < threads; t++) {
            for (uint64_t width = sub_leaves / 2; width >= 1; width /= 2) {
                uint64_t layer_first = width * threads + t * width; /* heap index */
                for (uint64_t k = 0; k < width; k++) {
                    uint64_t i = layer_first + k;
                    oracle_tip5_hash_pair(nodes + 5 * (2 * i), nodes + 5 * (2 * i + 1), nodes + 5 * i);
                }
            }
        }
        remaining = threads; /* remaining >>= subtree_height :207-209 */
    }
    for (uint64_t i = remaining - 1; i >= 1; i--) /* sequentially_fill_tree :216-222 */
        oracle_tip5_hash_pair(nodes + 5 * (2 * i), nodes + 5 * (2 * i + 1), nodes + 5 * i);
    return 0;
}

/* MmrAccumulator::peaks_from_leafs, mmr/mmr_accumulator.rs:96-115, restricted to what
 * sequential_frugal_root (merkle_tree.rs:299-309) needs. peaks: up to 64 digests. */
static uint64_t peaks_from_leafs(const uint64_t *leafs, uint64_t n, uint64_t peaks[64 * 5]) {
    uint64_t n_peaks = 0;
    for (uint64_t pair = 0; pair < n / 2; pair++) {
        uint64_t diagonal_idx = pair + 1;
        uint64_t right[5];
        oracle_tip5_hash_pair(leafs + 5 * (2 * pair), leafs + 5 * (2 * pair + 1), right);
        int merges = __builtin_ctzll(diagonal_idx);
        for (int k = 0; k < merges; k++) {
            n_peaks--;
            uint64_t merged[5];
            oracle_tip5_hash_pair(peaks + 5 * n_peaks, right, merged);
            memcpy(right, merged, sizeof(right));
        }
        memcpy(peaks + 5 * n_peaks, right, sizeof(right));
        n_peaks++;
    }
    if (n % 2 == 1) {
        memcpy(peaks + 5 * n_peaks, leafs + 5 * (n - 1), sizeof(uint64_t) * 5);
        n_peaks++;
    }
    return n_peaks;
}

/* MerkleTree::sequential_frugal_root, merkle_tree.rs:299-309 */
int oracle_merkle_sequential_frugal_root(const uint64_t *leafs, uint64_t n, uint64_t root[5]) {
    if (n == 0) return ORACLE_E_TOO_FEW_LEAFS;
    tip5_setup();
    uint64_t peaks[64 * 5];
    uint64_t n_peaks = peaks_from_leafs(leafs, n, peaks);
    if (n_peaks != 1) return ORACLE_E_INCORRECT_NUMBER_OF_LEAFS;
    memcpy(root, peaks, sizeof(uint64_t) * 5);
    return 0;
}

/* MerkleTree::par_frugal_root, merkle_tree.rs:332-364 */
int oracle_merkle_par_frugal_root(const uint64_t *leafs, uint64_t n, uint64_t root[5],
                                  int num_threads, uint64_t cutoff) {
    if (n == 0 || (n & (n - 1))) return ORACLE_E_INCORRECT_NUMBER_OF_LEAFS; /* :333-335 */
    tip5_setup();
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
    num_threads = 1;
#endif
    uint64_t threads = 1;
    while (threads * 2 <= (uint64_t)num_threads) threads *= 2;
    if (cutoff < 2) cutoff = 2;
    const uint64_t *cur = leafs;
    uint64_t cur_n = n;
    uint64_t *owned = NULL;
    int rc = 0;
    while (cur_n >= cutoff) {
        while (threads > cur_n / 2) threads /= 2;
        uint64_t chunk = cur_n / threads;
        uint64_t *next = (uint64_t *)malloc(sizeof(uint64_t) * 5 * threads);
#pragma omp parallel for schedule(static) num_threads((int)threads)
        for (uint64_t t = 0; t < threads; t++) {
            int r = oracle_merkle_sequential_frugal_root(cur + 5 * t * chunk, chunk, next + 5 * t);
            if (r) rc = r;
        }
        free(owned);
        owned = next;
        cur = next;
        cur_n = threads;
        if (rc) break;
    }
    if (!rc) rc = oracle_merkle_sequential_frugal_root(cur, cur_n, root);
    free(owned);
    return rc;
}

/* MmrAccumulator::peaks_from_leafs for any leaf count (mmr/mmr_accumulator.rs:96-115); returns the number
 * of peaks written (at most 64 digests). */
uint64_t oracle_mmr_peaks_from_leafs(const uint64_t *leafs, uint64_t n, uint64_t *peaks) {
    tip5_setup();
    return peaks_from_leafs(leafs, n, peaks);
}

/* bag_peaks, mmr/mmr_accumulator.rs:379-391: hash_10 of the BFieldCodec encoding of leaf_count (two 32-bit
 * limbs, bfield_codec.rs:122-128) padded with zeros, then folded from the last peak: acc = hash_pair(peak, acc) */
void oracle_mmr_bag_peaks(const uint64_t *peaks, uint64_t n_peaks, uint64_t leaf_count, uint64_t out[5]) {
    tip5_setup();
    uint64_t in[10] = {0};
    in[0] = bfe_new(leaf_count & 0xffffffffull);
    in[1] = bfe_new(leaf_count >> 32);
    uint64_t acc[5];
    oracle_tip5_hash_10(in, acc);
    for (uint64_t k = n_peaks; k-- > 0;) {
        uint64_t next[5];
        oracle_tip5_hash_pair(peaks + 5 * k, acc, next);
        memcpy(acc, next, sizeof(acc));
    }
    memcpy(out, acc, sizeof(acc));
}

/* MerkleTree::authentication_structure_node_indices, merkle_tree.rs:449-504: the siblings along every
 * leaf-to-root path that cannot be computed from the paths themselves, sorted by descending node index.
 * out needs room for n_indices * log2(num_leafs) entries; returns the count or a negative error
 * (-4 IncorrectNumberOfLeafs :467-469, -9 LeafIndexInvalid :487-489). */
static int cmp_u64_desc(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? 1 : x > y ? -1 : 0;
}
static int cmp_u64_asc(const void *a, const void *b) { return -cmp_u64_desc(a, b); }
int64_t oracle_auth_structure_node_indices(uint64_t num_leafs, const uint64_t *leaf_indices, uint64_t n_indices,
                                           uint64_t *out) {
    if (num_leafs == 0 || (num_leafs & (num_leafs - 1))) return ORACLE_E_INCORRECT_NUMBER_OF_LEAFS;
    uint64_t height = 0;
    while ((1ull << height) < num_leafs) height++;
    uint64_t cap = n_indices * height + 1;
    uint64_t *needed = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    uint64_t *computable = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    uint64_t nn = 0, nc = 0;
    for (uint64_t k = 0; k < n_indices; k++) {
        if (leaf_indices[k] >= num_leafs) {
            free(needed);
            free(computable);
            return -9;
        }
        uint64_t node = leaf_indices[k] + num_leafs;
        while (node > 1) {
            computable[nc++] = node;
            needed[nn++] = node ^ 1;
            node /= 2;
        }
    }
    qsort(needed, nn, sizeof(uint64_t), cmp_u64_desc);
    qsort(computable, nc, sizeof(uint64_t), cmp_u64_asc);
    int64_t count = 0;
    for (uint64_t k = 0; k < nn; k++) {
        if (k && needed[k] == needed[k - 1]) continue; /* set semantics */
        if (bsearch(&needed[k], computable, nc, sizeof(uint64_t), cmp_u64_asc)) continue;
        out[count++] = needed[k];
    }
    free(needed);
    free(computable);
    return count;
}

/* ------------------------------------------------------------------------------------------
 * Field helpers exported for tests
 * ---------------------------------------------------------------------------------------- */
uint64_t oracle_bfe_new(uint64_t v) { return bfe_new(v); }
uint64_t oracle_bfe_value(uint64_t raw) { return bfe_value(raw); }
uint64_t oracle_bfe_add(uint64_t a, uint64_t b) { return bfe_add(a, b); }
uint64_t oracle_bfe_sub(uint64_t a, uint64_t b) { return bfe_sub(a, b); }
uint64_t oracle_bfe_mul(uint64_t a, uint64_t b) { return bfe_mul(a, b); }
uint64_t oracle_bfe_mod_pow(uint64_t a, uint64_t e) { return bfe_mod_pow(a, e); }
uint64_t oracle_bfe_inverse_or_zero(uint64_t a) { return bfe_inverse_or_zero(a); }
uint64_t oracle_bfe_primitive_root_of_unity(uint64_t n) {
    if (n == 0) return bfe_new(1);
    return bfe_primitive_root_of_unity_log2(63 - __builtin_clzll(n));
}
void oracle_bfe_new_array(uint64_t *x, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) x[i] = bfe_new(x[i]);
}
void oracle_bfe_value_array(uint64_t *x, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) x[i] = bfe_value(x[i]);
}

/* Polynomial::evaluate (Horner) over BFE, used by the NTT==evaluation property test
 * (ntt.rs:562-579) and the coset tests (polynomial.rs:3645-3662). */
uint64_t oracle_poly_evaluate(const uint64_t *coeffs, uint64_t n_coeffs, uint64_t x_raw) {
    uint64_t acc = 0;
    for (uint64_t i = n_coeffs; i-- > 0;) acc = bfe_add(bfe_mul(acc, x_raw), coeffs[i]);
    return acc;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* rayon sizes its pool from the machine, not from OMP_NUM_THREADS (which torchrun sets to 1) */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- polynomial multiplication (next-wave checker) --------------------------------
 * XFieldElement * XFieldElement, x_field_element.rs:512-535 ([c,b,a] = coefficients, mod x^3 - x + 1). */
static void xfe_mul(const uint64_t l[3], const uint64_t r[3], uint64_t out[3]) {
    uint64_t c = l[0], b = l[1], a = l[2], f = r[0], e = r[1], d = r[2];
    uint64_t ae = bfe_mul(a, e), bd = bfe_mul(b, d), ad = bfe_mul(a, d);
    out[0] = bfe_sub(bfe_sub(bfe_mul(c, f), ae), bd);
    out[1] = bfe_add(bfe_add(bfe_sub(bfe_add(bfe_mul(b, f), bfe_mul(c, e)), ad), ae), bd);
    out[2] = bfe_add(bfe_add(bfe_add(bfe_mul(a, f), bfe_mul(b, e)), bfe_mul(c, d)), ad);
}

/* schoolbook product (Polynomial::naive_multiply, polynomial.rs `naive_multiply`): out has na+nb-1 elements */
void oracle_poly_naive_multiply(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint32_t w,
                                uint64_t *out) {
    if (na == 0 || nb == 0) return;
    memset(out, 0, (na + nb - 1) * w * sizeof(uint64_t));
    for (uint64_t i = 0; i < na; i++)
        for (uint64_t j = 0; j < nb; j++) {
            if (w == 1) {
                out[i + j] = bfe_add(out[i + j], bfe_mul(a[i], b[j]));
            } else {
                uint64_t t[3];
                xfe_mul(a + 3 * i, b + 3 * j, t);
                for (int k = 0; k < 3; k++) out[3 * (i + j) + k] = bfe_add(out[3 * (i + j) + k], t[k]);
            }
        }
}

/* Polynomial::fast_multiply, polynomial.rs:900-932: resize to the next power of two, ntt both,
 * Hadamard product, intt, truncate. */
int oracle_poly_fast_multiply(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint32_t w,
                              uint64_t *out) {
    if (na == 0 || nb == 0) return 0;
    uint64_t len = na + nb - 1, order = 1;
    while (order < len) order <<= 1;
    uint64_t *l = calloc(order * w, sizeof(uint64_t)), *r = calloc(order * w, sizeof(uint64_t));
    if (!l || !r) { free(l); free(r); return -6; }
    memcpy(l, a, na * w * sizeof(uint64_t));
    memcpy(r, b, nb * w * sizeof(uint64_t));
    int rc = oracle_ntt(l, order, w);
    if (!rc) rc = oracle_ntt(r, order, w);
    if (!rc) {
        for (uint64_t i = 0; i < order; i++) {
            if (w == 1) l[i] = bfe_mul(l[i], r[i]);
            else { uint64_t t[3]; xfe_mul(l + 3 * i, r + 3 * i, t); memcpy(l + 3 * i, t, sizeof(t)); }
        }
        rc = oracle_intt(l, order, w);
    }
    if (!rc) memcpy(out, l, len * w * sizeof(uint64_t));
    free(l);
    free(r);
    return rc;
}

/* ---- out-of-domain evaluation / coset extrapolation (next-wave checker) ---------------
 * Polynomial::evaluate, polynomial.rs:309-319 (Horner from the highest coefficient), for BFE (w = 1) or XFE
 * (w = 3) coefficients and a point of the same field. */
void oracle_poly_evaluate_w(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t w, const uint64_t *x, uint64_t *out) {
    uint64_t acc[3] = {0, 0, 0};
    for (uint64_t i = n_coeffs; i-- > 0;) {
        if (w == 1) {
            acc[0] = bfe_add(bfe_mul(acc[0], x[0]), coeffs[i]);
        } else {
            uint64_t t[3];
            xfe_mul(acc, x, t);
            for (int k = 0; k < 3; k++) acc[k] = bfe_add(t[k], coeffs[3 * i + k]);
        }
    }
    for (uint32_t k = 0; k < w; k++) out[k] = acc[k];
}

/* Polynomial::naive_coset_extrapolate applied per codeword (polynomial.rs:2135-2147, 2310-2331 without the
 * modular reduction, which does not change the values): intt, scale by offset^-1, evaluate in every point.
 * out[(codeword * n_points + point) * w ..] */
int oracle_batch_coset_extrapolate(uint64_t offset_raw, uint64_t n, const uint64_t *codewords, uint64_t n_codewords,
                                   uint32_t w, const uint64_t *points, uint64_t n_points, uint64_t *out) {
    if (n == 0 || (n & (n - 1))) return ORACLE_E_LEN_NOT_POW2;
    uint64_t *coeffs = (uint64_t *)malloc(sizeof(uint64_t) * n * w);
    int rc = 0;
    for (uint64_t c = 0; c < n_codewords && !rc; c++) {
        rc = oracle_coset_interpolate(codewords + c * n * w, n, w, offset_raw, coeffs);
        for (uint64_t p = 0; p < n_points && !rc; p++)
            oracle_poly_evaluate_w(coeffs, n, w, points + p * w, out + (c * n_points + p) * w);
    }
    free(coeffs);
    return rc;
}

/* Tip5::sample_indices, tip5/mod.rs:636-656 (squeeze = emit state[..RATE], then permute, :693-698) */
void oracle_tip5_sample_indices(uint64_t state[16], uint32_t upper_bound, uint64_t num_indices, uint32_t *out) {
    tip5_setup();
    uint64_t produced = 0, buffer[10];
    int next_in_buffer = 10;
    while (produced < num_indices) {
        if (next_in_buffer == 10) {
            memcpy(buffer, state, sizeof(buffer));
            oracle_tip5_permutation(state);
            next_in_buffer = 0;
        }
        uint64_t element = buffer[next_in_buffer++];
        if (bfe_value(element) != 0xFFFFFFFF00000000ull) out[produced++] = (uint32_t)bfe_value(element) % upper_bound;
    }
}

/* Polynomial::naive_divide, polynomial.rs:552-612 (BFieldElement): quotient (na - nb + 1 coefficients after
 * trimming) and remainder by long division from the leading coefficient.  The checker of clean_divide
 * (polynomial.rs:2358-2413), whose result is this quotient whenever the remainder is zero.
 * Returns the number of quotient coefficients, -1 for a zero divisor.  rem needs na words. */
int64_t oracle_poly_naive_divide(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint64_t *quot,
                                 uint64_t *rem) {
    while (na && a[na - 1] == 0) na--;
    while (nb && b[nb - 1] == 0) nb--;
    if (nb == 0) return -1;
    memcpy(rem, a, na * sizeof(uint64_t));
    if (na < nb) return 0;
    uint64_t lc_inv = oracle_bfe_inverse_or_zero(b[nb - 1]);
    uint64_t qd = na - nb;
    for (uint64_t k = 0; k <= qd; k++) {
        uint64_t top = na - 1 - k; /* current leading position of the remainder */
        uint64_t qc = bfe_mul(rem[top], lc_inv);
        quot[qd - k] = qc;
        if (qc != 0)
            for (uint64_t i = 0; i < nb; i++) rem[top - i] = bfe_sub(rem[top - i], bfe_mul(qc, b[nb - 1 - i]));
    }
    return (int64_t)(qd + 1);
}

/* Polynomial::reduce_by_ntt_friendly_modulus, polynomial.rs:1087-1148, statement by statement (w = 1 or 3).
 * out needs max(n, domain_length) elements; returns the number of elements of the result. */
int64_t oracle_poly_reduce_by_ntt_friendly_modulus(const uint64_t *coeffs, uint64_t n, uint32_t w,
                                                   const uint64_t *shift_ntt, uint64_t domain_length,
                                                   uint64_t tail_length, uint64_t *out) {
    if (domain_length == 0 || (domain_length & (domain_length - 1))) return ORACLE_E_LEN_NOT_POW2;
    uint64_t chunk_size = domain_length - tail_length;
    if (n < chunk_size + tail_length) { /* :1097-1099 */
        memcpy(out, coeffs, n * w * sizeof(uint64_t));
        return (int64_t)n;
    }
    uint64_t num_chunks = (n - (tail_length + chunk_size) + chunk_size - 1) / chunk_size;
    uint64_t range_start = num_chunks * chunk_size;
    uint64_t *window = (uint64_t *)calloc(domain_length * w, sizeof(uint64_t));
    uint64_t *product = (uint64_t *)calloc(domain_length * w, sizeof(uint64_t));
    uint64_t *next = (uint64_t *)calloc(domain_length * w, sizeof(uint64_t));
    if (range_start < n) memcpy(window, coeffs + range_start * w, (n - range_start) * w * sizeof(uint64_t));
    for (uint64_t ci = num_chunks; ci-- > 0;) {
        memset(product, 0, domain_length * w * sizeof(uint64_t));
        memcpy(product, window + tail_length * w, chunk_size * w * sizeof(uint64_t));
        oracle_ntt(product, domain_length, w);
        for (uint64_t i = 0; i < domain_length; i++) {
            if (w == 1) product[i] = bfe_mul(product[i], shift_ntt[i]);
            else { uint64_t t[3]; xfe_mul(product + 3 * i, shift_ntt + 3 * i, t); memcpy(product + 3 * i, t, sizeof(t)); }
        }
        oracle_intt(product, domain_length, w);
        memset(next, 0, domain_length * w * sizeof(uint64_t));
        memcpy(next + chunk_size * w, window, tail_length * w * sizeof(uint64_t));
        memcpy(next, coeffs + ci * chunk_size * w, chunk_size * w * sizeof(uint64_t));
        for (uint64_t i = 0; i < domain_length * w; i++) next[i] = bfe_sub(next[i], product[i]);
        uint64_t *tmp = window; window = next; next = tmp;
    }
    memcpy(out, window, domain_length * w * sizeof(uint64_t));
    free(window); free(product); free(next);
    return (int64_t)domain_length;
}
