/*
 * oracle/field.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of twenty-first's BFieldElement arithmetic.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.  The shipped path (twenty-first_b200/csrc) never includes it.
 *
 * Everything here works on the *raw Montgomery word* of a BFieldElement, exactly as the
 * reference stores it in memory (reference: twenty-first/src/math/b_field_element.rs:84-86).
 *
 * Parity status: PINNED against the reference's own known-answer tests
 * (tests/test_oracle_kat.py): b_field_element.rs:1478-1514, ntt.rs:397-469,511-560,
 * tip5/mod.rs:1145-1206,1294-1362,1525-1531.
 */
#ifndef TF21_ORACLE_FIELD_H
#define TF21_ORACLE_FIELD_H

#include <stdint.h>

typedef unsigned __int128 u128;

/* b_field_element.rs:225 */
#define BFE_P 0xffffffff00000001ULL
/* b_field_element.rs:229 -- 2^128 mod P */
#define BFE_R2 0xfffffffe00000001ULL

/* Montgomery reduction, b_field_element.rs:357-370 */
static inline uint64_t bfe_montyred(u128 x) {
    uint64_t xl = (uint64_t)x;
    uint64_t xh = (uint64_t)(x >> 64);
    uint64_t a = xl + (xl << 32);
    uint64_t e = a < xl; /* overflowing_add carry */
    uint64_t b = a - (a >> 32) - e;
    uint64_t r = xh - b;
    uint64_t c = xh < b; /* overflowing_sub borrow */
    return r - (1 + ~BFE_P) * c;
}

/* BFieldElement::new, b_field_element.rs:235-237 : canonical-or-not u64 -> raw word */
static inline uint64_t bfe_new(uint64_t v) { return bfe_montyred((u128)v * (u128)BFE_R2); }

/* BFieldElement::value / canonical_representation, b_field_element.rs:248, 334-336 */
static inline uint64_t bfe_value(uint64_t raw) { return bfe_montyred((u128)raw); }

/* impl Add, b_field_element.rs:711-732 : a + b = a - (p - b) */
static inline uint64_t bfe_add(uint64_t a, uint64_t b) {
    uint64_t pb = BFE_P - b;
    uint64_t x1 = a - pb;
    return (a < pb) ? x1 + BFE_P : x1;
}

/* impl Sub, b_field_element.rs:773-795 */
static inline uint64_t bfe_sub(uint64_t a, uint64_t b) {
    uint64_t x1 = a - b;
    uint64_t c1 = a < b;
    return x1 - (1 + ~BFE_P) * c1;
}

/* impl Mul, b_field_element.rs:755-762 */
static inline uint64_t bfe_mul(uint64_t a, uint64_t b) { return bfe_montyred((u128)a * (u128)b); }

/* mod_pow, b_field_element.rs:340-355 (square-and-multiply, MSB first) */
static inline uint64_t bfe_mod_pow(uint64_t base, uint64_t exp) {
    uint64_t acc = bfe_new(1);
    int bit_length = exp ? 64 - __builtin_clzll(exp) : 0;
    for (int i = 0; i < bit_length; i++) {
        acc = bfe_mul(acc, acc);
        if (exp & (1ULL << (bit_length - 1 - i))) acc = bfe_mul(acc, base);
    }
    return acc;
}

/* inverse, b_field_element.rs:254-284 computes x^(p-2) by an addition chain; the result
 * is the unique field inverse, so the plain Fermat power is an exact restatement.
 * inverse_or_zero maps 0 -> 0. */
static inline uint64_t bfe_inverse_or_zero(uint64_t x) {
    if (x == 0) return 0;
    return bfe_mod_pow(x, BFE_P - 2);
}

/* PRIMITIVE_ROOTS, b_field_element.rs:43-78, indexed by log2(n); canonical values. */
static const uint64_t BFE_PRIMITIVE_ROOTS_BY_LOG2[33] = {
    1ULL,
    18446744069414584320ULL,
    281474976710656ULL,
    18446744069397807105ULL,
    17293822564807737345ULL,
    70368744161280ULL,
    549755813888ULL,
    17870292113338400769ULL,
    13797081185216407910ULL,
    1803076106186727246ULL,
    11353340290879379826ULL,
    455906449640507599ULL,
    17492915097719143606ULL,
    1532612707718625687ULL,
    16207902636198568418ULL,
    17776499369601055404ULL,
    6115771955107415310ULL,
    12380578893860276750ULL,
    9306717745644682924ULL,
    18146160046829613826ULL,
    3511170319078647661ULL,
    17654865857378133588ULL,
    5416168637041100469ULL,
    16905767614792059275ULL,
    9713644485405565297ULL,
    5456943929260765144ULL,
    17096174751763063430ULL,
    1213594585890690845ULL,
    6414415596519834757ULL,
    16116352524544190054ULL,
    9123114210336311365ULL,
    4614640910117430873ULL,
    1753635133440165772ULL,
};

/* primitive_root_of_unity, b_field_element.rs:814-818; n must be 0, or 2^k with k<=32.
 * Returns the raw word. */
static inline uint64_t bfe_primitive_root_of_unity_log2(unsigned log2_n) {
    return bfe_new(BFE_PRIMITIVE_ROOTS_BY_LOG2[log2_n]);
}

#endif
