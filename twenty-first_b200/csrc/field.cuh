// field.cuh -- Goldilocks (p = 2^64 - 2^32 + 1) arithmetic for sm_100a.
//
// Replaces the scalar Montgomery arithmetic of the reference
// (twenty-first/src/math/b_field_element.rs:357-370 montyred, :711-732 Add, :773-795 Sub, :755-762 Mul).
// The kernels never use Montgomery multiplication: NTT is F_p-linear and Tip5's S-box exponent 7
// satisfies R^6 = 1, so plain mod-p arithmetic on the raw Montgomery words with canonical
// twiddles / raw round constants reproduces the reference bit for bit.
//
// Representation inside kernels: "weak" = any u64 (value mod p, possibly >= p).  Only the final
// store (and Tip5's LUT lanes) canonicalise to [0, p).
//
// Instruction budget (SASS, counted with cuobjdump): gl_add 7, gl_add_weak 5, gl_sub 5, gl_mul 18,
// gl_canon 5.
#pragma once
#include <cstdint>

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ull
#define GL_EPS 0xFFFFFFFFull /* 2^64 mod p */

#ifndef TF21_PRED_FIX
#define TF21_PRED_FIX 1  /* wrap corrections as predicated instructions (+4 % on the 2^20 batch) */
#endif
#ifndef TF21_PRED_MUL
#define TF21_PRED_MUL 1
#endif
#ifndef TF21_SHL_WIDE
#define TF21_SHL_WIDE 0  /* limbs of x << t in gl_shlc, tools/ab.sh on the 2^20 batch (3.54 / 3.64 / 3.67 / 3.54 / 3.67 ms for 0..4): 0: C shifts (ptxas splits them over both pipes), 1: two IMAD.WIDE with a 64-bit addend, 2: OR instead of the addend, 3: low limb as a narrow IMAD + funnel shifts, 4: one plain mul.wide + ALU */
#endif
#ifndef TF21_SUB_WIDE
#define TF21_SUB_WIDE 0  /* measured -3 %: the FMA pipe is as loaded as the ALU pipe */
#endif

#ifndef TF21_SHL_CARRY
#define TF21_SHL_CARRY 1  /* carry-form canonicalisation in the q == 0 fold and z0 * EPS on the ALU in the q == 2 fold: 2^20 batch 2.97 -> 2.92 ms */
#endif
#ifndef TF21_SHL_Q1_CARRY
#define TF21_SHL_Q1_CARRY 0  /* neutral: ptxas then forms z0 + z1 as LEA (ALU pipe) instead of IMAD.SHL + IMAD.IADD, the ALU count stays 7 */
#endif
#ifndef TF21_CANON_CARRY
#define TF21_CANON_CARRY 1  /* carry form of the canonicalisation: one ALU instruction less per use, 2^20 batch 3.02 -> 2.96 ms */
#endif
#ifndef TF21_MUL_WIDE_LOW
#define TF21_MUL_WIDE_LOW 1
#endif

#ifdef __CUDACC__

// 2^t as an operand ptxas cannot see through (a __constant__ may be rewritten by the host), so that
// x * 2^t stays an IMAD.WIDE on the FMA pipe instead of becoming shifts on the saturated ALU pipe.
__constant__ u32 c_gl_pow2[32] = {1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,
                                  1u << 8,  1u << 9,  1u << 10, 1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15,
                                  1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22, 1u << 23,
                                  1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

__device__ __forceinline__ u64 gl_pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// [0, 2^64) -> [0, p):  t = a - p borrows  <=>  a < p.  Only borrow-type flags are used: a `subc`
// that consumes the carry of an `add.cc` does NOT yield -carry (measured on sm_100a), so carry
// and borrow chains are never mixed in this file.
__device__ __forceinline__ u64 gl_canon(u64 a) {
    u32 lo, hi, m;
    asm("{\n\t"
        ".reg .u32 l2, h2;\n\t"
        "sub.cc.u32 l2,%3,1;\n\t"
        "subc.cc.u32 h2,%4,0xffffffff;\n\t"
        "subc.u32 %2,0,0;\n\t" /* m = (a < p) ? 0xffffffff : 0 */
        "lop3.b32 %0,l2,%3,%2,0xD8;\n\t" /* m ? lo : l2 */
        "lop3.b32 %1,h2,%4,%2,0xD8;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi), "=&r"(m)
        : "r"((u32)a), "r"((u32)(a >> 32)));
    return gl_pack(lo, hi);
}

// weak add: a, b any u64 with a + b < 2^64 + p (e.g. one of them canonical) -> any u64.
// One wrap correction: 2^64 = EPS (mod p); c*EPS is added as (c << 32) - c, which cannot wrap again.
__device__ __forceinline__ u64 gl_add_weak(u64 a, u64 b) {
    u32 lo, hi, m;
    asm("{\n\t"
        "add.cc.u32 %0,%3,%5;\n\t"
        "addc.cc.u32 %1,%4,%6;\n\t"
        "addc.u32 %2,0,0;\n\t" /* c = carry (0/1);  r += c * EPS = (c << 32) - c */
        "sub.cc.u32 %0,%0,%2;\n\t"
        "subc.u32 %1,%1,0;\n\t"
        "add.u32 %1,%1,%2;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi), "=&r"(m)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return gl_pack(lo, hi);
}

// sub: a any u64, b < p -> any u64; a, b both canonical -> canonical (a - b + p on borrow).
__device__ __forceinline__ u64 gl_sub(u64 a, u64 b) {
    u32 lo, hi, m;
    asm("{\n\t"
        "sub.cc.u32 %0,%3,%5;\n\t"
        "subc.cc.u32 %1,%4,%6;\n\t"
        "subc.u32 %2,0,0;\n\t" /* m = borrow ? 0xffffffff : 0 */
        "sub.cc.u32 %0,%0,%2;\n\t"
        "subc.u32 %1,%1,0;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi), "=&r"(m)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
    return gl_pack(lo, hi);
}

// canonical add: a, b < p -> a + b mod p in [0, p), computed as a - (p - b) like the reference
// (b_field_element.rs:716-731).
__device__ __forceinline__ u64 gl_add(u64 a, u64 b) { return gl_sub(a, GL_P - b); }

// 64x64 -> 128 bit product as four 32-bit limbs. ptxas fuses the mad.lo.cc/madc.hi pairs into
// IMAD.WIDE.U32 with carry-out: 4 IMAD.WIDE + 4 moves.
__device__ __forceinline__ void gl_mul128(u64 a, u64 b, u32 &r0, u32 &r1, u32 &r2, u32 &r3) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t"
        "mul.lo.u32 %0,%4,%6;\n\t"
        "mul.hi.u32 %1,%4,%6;\n\t"
        "mad.lo.cc.u32 %1,%4,%7,%1;\n\t"
        "madc.hi.u32 %2,%4,%7,0;\n\t"
        "mad.lo.cc.u32 %1,%5,%6,%1;\n\t"
        "madc.hi.cc.u32 %2,%5,%6,%2;\n\t"
        "addc.u32 %3,0,0;\n\t"
        "mad.lo.cc.u32 %2,%5,%7,%2;\n\t"
        "madc.hi.u32 %3,%5,%7,%3;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

// Solinas reduction of r0 + r1 2^32 + r2 2^64 + r3 2^96 using 2^64 = 2^32 - 1, 2^96 = -1 (mod p)
// -> any u64.  x0 - r3 (one borrow fix) + r2 * EPS (one carry fix); neither fix can wrap twice.
__device__ __forceinline__ u64 gl_reduce128(u32 r0, u32 r1, u32 r2, u32 r3) {
    u32 lo, hi, m;
    asm("{\n\t"
        "sub.cc.u32  %0, %3, %6;\n\t"
        "subc.cc.u32 %1, %4, 0;\n\t"
        "subc.u32    %2, 0, 0;\n\t"
        "sub.cc.u32  %0, %0, %2;\n\t"
        "subc.u32    %1, %1, 0;\n\t"
        "mad.lo.cc.u32  %0, %5, 0xffffffff, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, 0xffffffff, %1;\n\t"
        "addc.u32    %2, 0, 0;\n\t" /* c = carry;  r += c * EPS = (c << 32) - c */
        "sub.cc.u32  %0, %0, %2;\n\t"
        "subc.u32    %1, %1, 0;\n\t"
        "add.u32     %1, %1, %2;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi), "=&r"(m)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return gl_pack(lo, hi);
}

// Same reduction with the two wrap corrections as predicated instructions (see gl_addp below):
// (r1:r0) + ~r3 + 1 in the carry domain, "no carry" = borrow -> -EPS; then + r2 * EPS, carry -> +EPS.
__device__ __forceinline__ u64 gl_reduce128p(u32 r0, u32 r1, u32 r2, u32 r3) {
    u64 out;
    asm("{\n\t.reg .u32 lo,hi,c,d,n3; .reg .pred p; .reg .u64 v;\n\t"
        "not.b32 n3,%4;\n\t"
        "add.cc.u32 d,0xffffffff,1;\n\t"
        "addc.cc.u32 lo,%1,n3;\n\t"
        "addc.cc.u32 hi,%2,0xffffffff;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.ne.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLR1%=;\n\t"
        "sub.u64 v,v,0xffffffff;\n\t"
        "GLR1%=:\n\t"
        "mov.b64 {lo,hi},v;\n\t"
        "mad.lo.cc.u32 lo,%3,0xffffffff,lo;\n\t"
        "madc.hi.cc.u32 hi,%3,0xffffffff,hi;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLR2%=;\n\t"
        "add.u64 v,v,0xffffffff;\n\t"
        "GLR2%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(out)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return out;
}

// mask-style reduction regardless of TF21_PRED_MUL: two registers less pressure than the predicated
// form inside the 1024-point column kernel (which otherwise spills)
__device__ __forceinline__ u64 gl_mul_mask(u64 a, u64 b) {
    u32 r0, r1, r2, r3;
    gl_mul128(a, b, r0, r1, r2, r3);
    return gl_reduce128(r0, r1, r2, r3);
}

// a * b mod p; a, b any u64 -> any u64
__device__ __forceinline__ u64 gl_mul(u64 a, u64 b) {
    u32 r0, r1, r2, r3;
    gl_mul128(a, b, r0, r1, r2, r3);
#if TF21_PRED_MUL
    return gl_reduce128p(r0, r1, r2, r3);
#else
    return gl_reduce128(r0, r1, r2, r3);
#endif
}

// ---- ALU-only Solinas folds for FMA-pipe-bound code (Tip5) --------------------------------------------
// In the Tip5 round the FMA-heavy pipe is the busiest (48 modular products = 192 IMAD.WIDE at one warp
// instruction per 5.3 cycles, ~78 % busy) while the ALU pipe is at ~45 %.  The r2 * EPS term of the folds above
// costs one IMAD + one IMAD.HI (7.3 FMA-pipe cycles); here it is built on the ALU as the 64-bit difference
// (r2 << 32) - r2 = (-r2, r2 - (r2 != 0)) and added with carry: 4 ALU instructions, no multiply.
#ifndef TF21_EPS_NOT
#define TF21_EPS_NOT 0
#endif
#if TF21_EPS_NOT  /* (~r2 + 1, r2 + 0xffffffff + carry): one more LOP3, no negate on the FMA pipe */
#define GL_EPS_TERM(R) "not.b32 n2," R ";\n\tadd.cc.u32 tl,n2,1;\n\taddc.u32 th," R ",0xffffffff;\n\t"
#else
#define GL_EPS_TERM(R) "sub.cc.u32 tl,0," R ";\n\tsubc.u32 th," R ",0;\n\t"
#endif
__device__ __forceinline__ u64 gl_reduce96a(u64 x0, u32 x1) {  // x0 + x1 2^64 -> any u64
    u64 out;
    asm("{\n\t.reg .u32 lo,hi,c,tl,th,n2; .reg .pred p; .reg .u64 v;\n\t"
        "mov.b64 {lo,hi},%1;\n\t"
        GL_EPS_TERM("%2")
        "add.cc.u32 lo,lo,tl;\n\t"
        "addc.cc.u32 hi,hi,th;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLNA%=;\n\t"
        "add.u64 v,v,0xffffffff;\n\t"
        "GLNA%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(out)
        : "l"(x0), "r"(x1));
    return out;
}

__device__ __forceinline__ u64 gl_reduce128a(u32 r0, u32 r1, u32 r2, u32 r3) {
    u64 out;
    asm("{\n\t.reg .u32 lo,hi,c,d,n3,tl,th,n2; .reg .pred p; .reg .u64 v;\n\t"
        "not.b32 n3,%4;\n\t"
        "add.cc.u32 d,0xffffffff,1;\n\t"
        "addc.cc.u32 lo,%1,n3;\n\t"
        "addc.cc.u32 hi,%2,0xffffffff;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.ne.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLRA1%=;\n\t"
        "sub.u64 v,v,0xffffffff;\n\t"
        "GLRA1%=:\n\t"
        "mov.b64 {lo,hi},v;\n\t"
        GL_EPS_TERM("%3")
        "add.cc.u32 lo,lo,tl;\n\t"
        "addc.cc.u32 hi,hi,th;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLRA2%=;\n\t"
        "add.u64 v,v,0xffffffff;\n\t"
        "GLRA2%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(out)
        : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return out;
}

// 64x64 -> 128 with the low product as one mul.wide (ptxas keeps mul.lo + mul.hi as IMAD + IMAD.HI)
__device__ __forceinline__ void gl_mul128w(u64 a, u64 b, u32 &r0, u32 &r1, u32 &r2, u32 &r3) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u64 t;\n\t"
        "mul.wide.u32 t,%4,%6;\n\t"
        "mov.b64 {%0,%1},t;\n\t"
        "mad.lo.cc.u32 %1,%4,%7,%1;\n\t"
        "madc.hi.u32 %2,%4,%7,0;\n\t"
        "mad.lo.cc.u32 %1,%5,%6,%1;\n\t"
        "madc.hi.cc.u32 %2,%5,%6,%2;\n\t"
        "addc.u32 %3,0,0;\n\t"
        "mad.lo.cc.u32 %2,%5,%7,%2;\n\t"
        "madc.hi.u32 %3,%5,%7,%3;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

// 64x64 -> 128 as four plain wide multiplies (4.2 cycles each, against 5.3 for the multiply-add with a 64-bit
// addend) + 6 full-rate additions (tools/ubench3.cu)
__device__ __forceinline__ void gl_mul128_nw(u64 a, u64 b, u32 &r0, u32 &r1, u32 &r2, u32 &r3) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u64 p00,p01,p10,p11; .reg .u32 l1,h1,l2,h2;\n\t"
        "mul.wide.u32 p00,%4,%6;\n\t"
        "mul.wide.u32 p01,%4,%7;\n\t"
        "mul.wide.u32 p10,%5,%6;\n\t"
        "mul.wide.u32 p11,%5,%7;\n\t"
        "mov.b64 {%0,%1},p00;\n\t"
        "mov.b64 {%2,%3},p11;\n\t"
        "mov.b64 {l1,h1},p01;\n\t"
        "mov.b64 {l2,h2},p10;\n\t"
        "add.cc.u32 %1,%1,l1;\n\t"
        "addc.cc.u32 %2,%2,h1;\n\t"
        "addc.u32 %3,%3,0;\n\t"
        "add.cc.u32 %1,%1,l2;\n\t"
        "addc.cc.u32 %2,%2,h2;\n\t"
        "addc.u32 %3,%3,0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
__device__ __forceinline__ u64 gl_mul_nw(u64 a, u64 b) {
    u32 r0, r1, r2, r3;
    gl_mul128_nw(a, b, r0, r1, r2, r3);
    return gl_reduce128p(r0, r1, r2, r3);
}

// x^2 mod p with three 32x32 products instead of four: x^2 = x0^2 + 2 x0 x1 2^32 + x1^2 2^64; the doubled cross
// product is formed by adding it twice into the 96-bit window (x0^2 >> 32 | x1^2 << 32) -- 3 wide multiplies
// (quarter rate: 4.2 cycles per warp instruction) + 6 full-rate additions instead of 1 + 3 multiply-adds.
__device__ __forceinline__ u64 gl_sqr(u64 x) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32);
    u32 r0, r1, r2, r3;
    asm("{\n\t.reg .u64 p00,p01,p11; .reg .u32 m0,m1;\n\t"
        "mul.wide.u32 p00,%4,%4;\n\t"
        "mul.wide.u32 p01,%4,%5;\n\t"
        "mul.wide.u32 p11,%5,%5;\n\t"
        "mov.b64 {%0,%1},p00;\n\t"
        "mov.b64 {%2,%3},p11;\n\t"
        "mov.b64 {m0,m1},p01;\n\t"
        "add.cc.u32 %1,%1,m0;\n\t"
        "addc.cc.u32 %2,%2,m1;\n\t"
        "addc.u32 %3,%3,0;\n\t"
        "add.cc.u32 %1,%1,m0;\n\t"
        "addc.cc.u32 %2,%2,m1;\n\t"
        "addc.u32 %3,%3,0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(x0), "r"(x1));
    return gl_reduce128p(r0, r1, r2, r3);
}

// a * b mod p for Tip5's S-box: same value as gl_mul, reduction on the ALU pipe
__device__ __forceinline__ u64 gl_mul_alu(u64 a, u64 b) {
    u32 r0, r1, r2, r3;
#if TF21_MUL_WIDE_LOW
    gl_mul128w(a, b, r0, r1, r2, r3);
#else
    gl_mul128(a, b, r0, r1, r2, r3);
#endif
    return gl_reduce128a(r0, r1, r2, r3);
}

// canonical product
__device__ __forceinline__ u64 gl_mulc(u64 a, u64 b) { return gl_canon(gl_mul(a, b)); }

// 96-bit value x0 + x1 2^64 (x1 < 2^32) -> any u64
__device__ __forceinline__ u64 gl_reduce96(u64 x0, u32 x1) {
#if TF21_PRED_FIX
    u64 out;
    asm("{\n\t.reg .u32 lo,hi,c; .reg .pred p; .reg .u64 v;\n\t"
        "mov.b64 {lo,hi},%1;\n\t"
        "mad.lo.cc.u32 lo,%2,0xffffffff,lo;\n\t"
        "madc.hi.cc.u32 hi,%2,0xffffffff,hi;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLN%=;\n\t"
        "add.u64 v,v,0xffffffff;\n\t"
        "GLN%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(out)
        : "l"(x0), "r"(x1));
    return out;
#endif
    u32 lo, hi, m;
    asm("{\n\t"
        "mad.lo.cc.u32  %0, %5, 0xffffffff, %3;\n\t"
        "madc.hi.cc.u32 %1, %5, 0xffffffff, %4;\n\t"
        "addc.u32    %2, 0, 0;\n\t"
        "sub.cc.u32  %0, %0, %2;\n\t"
        "subc.u32    %1, %1, 0;\n\t"
        "add.u32     %1, %1, %2;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi), "=&r"(m)
        : "r"((u32)x0), "r"((u32)(x0 >> 32)), "r"(x1));
    return gl_pack(lo, hi);
}

// x * 2^S mod p for a compile-time (after unrolling) 0 <= S < 96; x < p -> result < p.
// Every root of unity of order <= 64 is a power of two (omega_64 = 2^39, b_field_element.rs:43-78),
// so the butterflies of a 32-point sub-transform need no 64x64 multiply: shift, then fold with
// 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 (mod p).
__device__ __forceinline__ u64 gl_mul_pow2(u64 x, const int S) {
    if (S == 0) return x;
    const int q = S >> 5, t = S & 31;
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32);
    u32 y0, y1, y2;
    if (t == 0) {
        y0 = x0;
        y1 = x1;
        y2 = 0;
    } else {
        y0 = x0 << t;
        y1 = __funnelshift_l(x0, x1, t);
        y2 = x1 >> (32 - t);
    }
    if (q == 0) {  // y0 + y1 2^32 + y2 2^64
        return gl_canon(gl_reduce96(gl_pack(y0, y1), y2));
    } else if (q == 1) {  // y0 2^32 + y1 2^64 + y2 2^96 = (y0 << 32) + y1 EPS - y2
        u64 a = (u64)y1 * GL_EPS;  // <= (2^32-1)^2 < p
        if (t != 0) a = gl_sub(a, (u64)y2);
        return gl_add(a, (u64)y0 << 32);  // (y0 << 32) <= p - 1
    } else {  // y0 2^64 + y1 2^96 + y2 2^128 = y0 EPS - (y1 + y2 2^32)
        u64 a = (u64)y0 * GL_EPS;
        return gl_sub(a, gl_pack(y1, y2));  // y2 < 2^31 => operand < p
    }
}

// ---- second-generation primitives: integer ALU / FMA-pipe balanced ---------------------------------
// Measured on B200 (tools/ubench.cu): IADD3/LOP3/SHF/SEL/ISETP issue at one warp instruction per
// 2 cycles per SM sub-partition on the ALU pipe; IMAD (32-bit) the same on the FMA pipe;
// IMAD.WIDE / IMAD.HI one per 5.3 cycles on the FMA pipe; both pipes run concurrently.  The
// butterfly-heavy kernels are ALU bound, so the wrap corrections below are expressed as one
// IMAD.WIDE (c * EPS + s) instead of two or three ALU instructions.

// s + c * EPS (mod 2^64), c in {0,1}.  Compiles to a single IMAD.WIDE.U32.
__device__ __forceinline__ u64 gl_fix(u64 s, u32 c) { return s + (u64)c * GL_EPS; }

// ---- wrap corrections as predicated instructions ---------------------------------------------------
// Written with an explicit branch around the 64-bit correction: ptxas folds the materialised carry back
// into the carry predicate of the IADD3.X and if-converts the two-instruction body, so a lazy add is
// IADD3, IADD3.X, @P IADD3, @P IADD3.X -- no SEL, no IMAD.WIDE.  Subtraction is a + ~b + 1 so that only
// add-with-carry instructions appear (borrow and carry flags are never mixed, see gl_canon).
#ifndef TF21_ADDFIX_WIDE
#define TF21_ADDFIX_WIDE 0  /* carry correction of the lazy addition as one predicated IMAD.WIDE.U32 (one * EPS + v) */
#endif
__device__ __forceinline__ u64 gl_addp_wide(u64 a, u64 t, u32 one) {  // a any, t <= p -> any
    u64 s;
    asm("{\n\t.reg .u32 c,lo,hi; .reg .pred p; .reg .u64 v;\n\t"
        "add.cc.u64 v,%1,%2;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "@p bra GLAW%=;\n\t"
        "mov.b64 {lo,hi},v;\n\t"
        "mad.lo.cc.u32 lo,%3,0xffffffff,lo;\n\t"
        "madc.hi.u32 hi,%3,0xffffffff,hi;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "GLAW%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(s)
        : "l"(a), "l"(t), "r"(one));
    return s;
}
__device__ __forceinline__ u64 gl_addp(u64 a, u64 t) {  // a any, t <= p -> any
    u64 s;
    asm("{\n\t.reg .u32 c; .reg .pred p; .reg .u64 v;\n\t"
        "add.cc.u64 v,%1,%2;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "@p bra GLA%=;\n\t"
        "add.u64 v,v,0xffffffff;\n\t"
        "GLA%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(s)
        : "l"(a), "l"(t));
    return s;
}
#ifndef TF21_SUB_MADC
#define TF21_SUB_MADC 1  /* high word of the borrow correction as IMAD.X (FMA pipe) instead of IADD3.X (ALU pipe) */
#endif
// an operand ptxas cannot fold (a __constant__ may be rewritten by the host): `one * 0xffffffff + hi + carry` stays a
// multiply-add with carry-in, i.e. IMAD.X on the FMA-heavy pipe -- the butterfly-heavy kernels are ALU-pipe bound
// (86 % against 55 %), so every addition that does not need a carry-OUT is worth moving
__constant__ u32 c_gl_one = 1u;

// `one`: a register holding 1 that ptxas cannot see through (kernels at the register limit load it once from global
// memory -- the constant-bank default is re-read before every use)
__device__ __forceinline__ u64 gl_subp(u64 a, u64 t, u32 one = c_gl_one) {  // a any, t <= p -> any; both < p -> result < p
    u64 s;
#if TF21_SUB_MADC
    asm("{\n\t.reg .u32 lo,hi,c,d,n0,n1; .reg .pred p;\n\t"
        "not.b32 n0,%3;\n\t"
        "not.b32 n1,%4;\n\t"
        "add.cc.u32 d,0xffffffff,1;\n\t"
        "addc.cc.u32 lo,%1,n0;\n\t"
        "addc.cc.u32 hi,%2,n1;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.ne.u32 p,c,0;\n\t"          /* no carry = borrow: subtract EPS, i.e. (lo + 1, hi - 1 + carry) */
        "@p bra GLSM%=;\n\t"
        "add.cc.u32 lo,lo,1;\n\t"
        "madc.lo.u32 hi,%5,0xffffffff,hi;\n\t"
        "GLSM%=:\n\t"
        "mov.b64 %0,{lo,hi};\n\t"
        "}"
        : "=l"(s)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)t), "r"((u32)(t >> 32)), "r"(one));
    return s;
#endif
    asm("{\n\t.reg .u32 lo,hi,c,d,n0,n1; .reg .pred p; .reg .u64 v;\n\t"
        "not.b32 n0,%3;\n\t"
        "not.b32 n1,%4;\n\t"
        "add.cc.u32 d,0xffffffff,1;\n\t"
        "addc.cc.u32 lo,%1,n0;\n\t"
        "addc.cc.u32 hi,%2,n1;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.ne.u32 p,c,0;\n\t"
        "mov.b64 v,{lo,hi};\n\t"
        "@p bra GLS%=;\n\t"
        "sub.u64 v,v,0xffffffff;\n\t"
        "GLS%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(s)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)t), "r"((u32)(t >> 32)));
    return s;
}
// carry form: x >= p  <=>  x + EPS carries out of 64 bits; then x - p = (x + EPS) mod 2^64.  Two adds with carry
// + a predicated 64-bit move instead of two compares + two predicated adds (one ALU instruction less).
__device__ __forceinline__ u64 gl_canonc(u64 x) {
    u64 s;
    asm("{\n\t.reg .u32 c; .reg .pred p; .reg .u64 v, w;\n\t"
        "add.cc.u64 w,%1,0xffffffff;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "mov.b64 v,%1;\n\t"
        "@p bra GLCC%=;\n\t"
        "mov.b64 v,w;\n\t"
        "GLCC%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(s)
        : "l"(x));
    return s;
}
#ifndef TF21_CANON_WIDE
#define TF21_CANON_WIDE 0  /* bit 0: canonicalisation of the trivial butterflies, bit 1: of the stored outputs -- x + EPS with its carry-out as ONE IMAD.WIDE.U32 (FMA-heavy pipe) instead of IADD3 + IADD3.X (ALU pipe) */
#endif
// x >= p  <=>  x + EPS carries; the sum is one * EPS + x on the FMA-heavy pipe (`one`: see gl_subp)
__device__ __forceinline__ u64 gl_canon_wide(u64 x, u32 one) {
    u64 s;
    asm("{\n\t.reg .u32 lo,hi,wl,wh,c; .reg .pred p; .reg .u64 v;\n\t"
        "mov.b64 {lo,hi},%1;\n\t"
        "mad.lo.cc.u32 wl,%2,0xffffffff,lo;\n\t"
        "madc.hi.cc.u32 wh,%2,0xffffffff,hi;\n\t"
        "addc.u32 c,0,0;\n\t"
        "setp.eq.u32 p,c,0;\n\t"
        "mov.b64 v,%1;\n\t"
        "@p bra GLCW%=;\n\t"
        "mov.b64 v,{wl,wh};\n\t"
        "GLCW%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(s)
        : "l"(x), "r"(one));
    return s;
}
__device__ __forceinline__ u64 gl_canonp(u64 x) {  // any -> [0, p)
#if TF21_CANON_CARRY
    return gl_canonc(x);
#endif
    u64 s;
    asm("{\n\t.reg .u32 lo,hi; .reg .pred p1,p2; .reg .u64 v;\n\t"
        "mov.b64 {lo,hi},%1;\n\t"
        "mov.b64 v,%1;\n\t"
        "setp.eq.u32 p1,hi,0xffffffff;\n\t"
        "setp.ne.and.u32 p2,lo,0,p1;\n\t"
        "@!p2 bra GLC%=;\n\t"
        "add.u64 v,v,0xffffffff;\n\t"
        "GLC%=:\n\t"
        "mov.b64 %0,v;\n\t"
        "}"
        : "=l"(s)
        : "l"(x));
    return s;
}

// lazy sub for the butterflies: a any u64, t <= p -> any u64.  With TF21_SUB_WIDE the wrap correction
// d - bw * EPS = (lo, hi - bw) + bw is one signed IMAD.WIDE (m * m + .., m = -bw) instead of two ALU ops.
__device__ __forceinline__ u64 gl_subl(u64 a, u64 t, u32 one = c_gl_one) {
#if TF21_PRED_FIX
    return gl_subp(a, t, one);
#elif TF21_SUB_WIDE
    u32 lo, hi, m;
    asm("sub.cc.u32 %0,%3,%5;\n\tsubc.cc.u32 %1,%4,%6;\n\tsubc.u32 %2,0,0;"
        : "=r"(lo), "=r"(hi), "=r"(m)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)t), "r"((u32)(t >> 32)));
    const long long d = (long long)gl_pack(lo, hi + m);
    return (u64)(d + (long long)(int)m * (long long)(int)m);
#else
    return gl_sub(a, t);
#endif
}

// lazy add: a any u64, t <= p  ->  any u64 (value a + t mod p).  3 ALU + 1 IMAD.WIDE.
__device__ __forceinline__ u64 gl_addl(u64 a, u64 t) {
#if TF21_PRED_FIX
    return gl_addp(a, t);
#endif
    u32 lo, hi, c;
    asm("add.cc.u32 %0,%3,%5;\n\taddc.cc.u32 %1,%4,%6;\n\taddc.u32 %2,0,0;"
        : "=r"(lo), "=r"(hi), "=r"(c)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)t), "r"((u32)(t >> 32)));
    return gl_fix(gl_pack(lo, hi), c);
}

// any u64 -> [0, p):  x >= p  <=>  hi == 0xffffffff && lo != 0.  2 ISETP + SEL + 1 IMAD.WIDE.
__device__ __forceinline__ u64 gl_canonw(u64 x) {
#if TF21_PRED_FIX
    return gl_canonp(x);
#endif
    u32 f;
    asm("{\n\t.reg .pred p1, p2;\n\t"
        "setp.eq.u32 p1, %2, 0xffffffff;\n\t"
        "setp.ne.and.u32 p2, %1, 0, p1;\n\t"
        "selp.u32 %0, 1, 0, p2;\n\t}"
        : "=r"(f)
        : "r"((u32)x), "r"((u32)(x >> 32)));
    return gl_fix(x, f);
}

// x * 2^S mod p with a CANONICAL result; x any u64; compile-time 0 < S < 96 with S % 32 != 0
// (every shift twiddle of a 32- or 64-point transform: S is a multiple of 3).
// z = x << (S % 32) as three 32-bit limbs, then fold by 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32.
// LAZY0: for S < 32 the result is only folded (any u64, < 2^64 - 2^32 unless its high word is 0xffffffff): the
// optimistic 32-point transform of ntt_fast.cuh checks the high words of a whole stage at once instead of
// canonicalising every operand.
template <int S, int V = TF21_SHL_WIDE, bool LAZY0 = false>
__device__ __forceinline__ u64 gl_shlc(u64 x, u32 one = c_gl_one) {
    static_assert(S > 0 && S < 96 && (S & 31) != 0, "shift twiddle out of range");
    constexpr int q = S >> 5, t = S & 31;
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32);
    u32 z0, z1, z2;
    if constexpr (V == 3) {
        // low limb on the FMA pipe (narrow IMAD, 2 cycles), the funnel shifts on the ALU pipe
        z0 = x0 * c_gl_pow2[t];
        z1 = __funnelshift_l(x0, x1, t);
        z2 = x1 >> (32 - t);
    } else if constexpr (V == 4) {
        // one plain wide multiply (no addend, 4.2 cycles) for (z1:z2) of the high word, ALU for the rest
        const u64 p1 = (u64)x1 * c_gl_pow2[t];
        z0 = x0 << t;
        z1 = (u32)p1 | (x0 >> (32 - t));
        z2 = (u32)(p1 >> 32);
    } else if constexpr (V == 1 || V == 2) {
        const u32 mt = c_gl_pow2[t];
        const u64 p0 = (u64)x0 * mt;
        const u64 p1 = (u64)x1 * mt;
        if constexpr (V == 2) {
            // hi(p0) < 2^t and the low t bits of lo(p1) are zero: OR instead of a 64-bit addend
            z0 = (u32)p0, z1 = (u32)p1 | (u32)(p0 >> 32), z2 = (u32)(p1 >> 32);
        } else {
            const u64 p1a = p1 + (p0 >> 32);
            z0 = (u32)p0, z1 = (u32)p1a, z2 = (u32)(p1a >> 32);
        }
    } else {
        z0 = x0 << t;
        z1 = __funnelshift_l(x0, x1, t);
        z2 = x1 >> (32 - t);  // < 2^31
    }
    if constexpr (q == 0 && LAZY0) {
        return gl_reduce96(gl_pack(z0, z1), z2);
    } else if constexpr (q == 0) {
        // (z1:z0) + z2 * EPS, z2 * EPS < 2^63: at most one wrap; carry and "r >= p" are exclusive
        u32 lo, hi, c;
        asm("mad.lo.cc.u32 %0,%5,0xffffffff,%3;\n\tmadc.hi.cc.u32 %1,%5,0xffffffff,%4;\n\taddc.u32 %2,0,0;"
            : "=r"(lo), "=r"(hi), "=r"(c)
            : "r"(z0), "r"(z1), "r"(z2));
#if TF21_SHL_CARRY
        // carry form: with s = (hi:lo) and c the carry of the fold, the result is s + EPS (mod 2^64) when c is set
        // (s < 2^63, no second carry) or when s + EPS carries (s >= p), else s: two adds, one predicate OR, a
        // predicated move -- two compares less than the form below
        u64 r;
        asm("{\n\t.reg .pred p1,p2; .reg .u32 k; .reg .u64 v,w;\n\t"
            "mov.b64 v,{%1,%2};\n\t"
            "add.cc.u64 w,v,0xffffffff;\n\t"
            "addc.u32 k,0,0;\n\t"
            "setp.ne.u32 p1,k,0;\n\t"
            "setp.ne.or.u32 p2,%3,0,p1;\n\t"
            "@!p2 bra GLQC%=;\n\t"
            "mov.b64 v,w;\n\t"
            "GLQC%=:\n\t"
            "mov.b64 %0,v;\n\t"
            "}"
            : "=l"(r)
            : "r"(lo), "r"(hi), "r"(c));
        return r;
#elif TF21_PRED_FIX
        u64 r;
        asm("{\n\t.reg .pred p1,p2,p3; .reg .u64 v;\n\t"
            "mov.b64 v,{%1,%2};\n\t"
            "setp.eq.u32 p1,%2,0xffffffff;\n\t"
            "setp.ne.and.u32 p2,%1,0,p1;\n\t"
            "setp.ne.or.u32 p3,%3,0,p2;\n\t"
            "@!p3 bra GLQ%=;\n\t"
            "add.u64 v,v,0xffffffff;\n\t"
            "GLQ%=:\n\t"
            "mov.b64 %0,v;\n\t"
            "}"
            : "=l"(r)
            : "r"(lo), "r"(hi), "r"(c));
        return r;
#else
        const u32 f = c | ((hi == 0xffffffffu && lo != 0) ? 1u : 0u);
        return gl_fix(gl_pack(lo, hi), f);
#endif
    } else if constexpr (q == 1) {
        // z * 2^32 = (z0 + z1) 2^32 - (z1 + z2):  T1 = (s : -carry) < p,  T2 = z1 + z2 < 2^33
#if TF21_SHL_Q1_CARRY
        // the carry of z0 + z1 from the adder itself (IADD3 + IMAD.X) instead of a compare + select (ISETP + SEL): one
        // ALU-pipe instruction less per shift by 32..63 bits
        u32 s, c;
        asm("add.cc.u32 %0,%2,%3;\n\taddc.u32 %1,0,0;" : "=r"(s), "=r"(c) : "r"(z0), "r"(z1));
#else
        const u32 s = z0 + z1;
        const u32 c = (s < z0) ? 1u : 0u;
#endif
        const u64 T1 = gl_pack(0u - c, s);
        const u64 T2 = (u64)z1 + (u64)z2;
        return gl_subl(T1, T2, one);
    } else {
        // z * 2^64 = z0 * EPS - (z2:z1):  z0 * EPS <= (2^32-1)^2 < p, (z2:z1) < 2^63
#if TF21_SHL_CARRY
        // z0 * EPS = (z0 << 32) - z0 as a 64-bit difference on the ALU instead of a wide multiply
        u32 alo, ahi;
        asm("sub.cc.u32 %0,0,%2;\n\tsubc.u32 %1,%2,0;" : "=r"(alo), "=r"(ahi) : "r"(z0));
        const u64 a = gl_pack(alo, ahi);
#else
        const u64 a = (u64)z0 * GL_EPS;
#endif
        return gl_subl(a, gl_pack(z1, z2), one);
    }
}

#endif  // __CUDACC__

// ---- host-side helpers (table construction only) -----------------------------------------------
static inline u64 hgl_mul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % GL_P); }
static inline u64 hgl_pow(u64 base, u64 e) {
    u64 acc = 1;
    base %= GL_P;
    while (e) {
        if (e & 1) acc = hgl_mul(acc, base);
        base = hgl_mul(base, base);
        e >>= 1;
    }
    return acc;
}
static inline u64 hgl_inv(u64 a) { return hgl_pow(a, GL_P - 2); }
// raw Montgomery word -> canonical value: v = raw * 2^-64 mod p
static inline u64 hgl_from_raw(u64 raw) { return hgl_mul(raw % GL_P, hgl_inv(GL_EPS)); }
static inline u64 hgl_to_raw(u64 v) { return hgl_mul(v % GL_P, GL_EPS); }
// omega_n = 7^((p-1)/n), n = 2^log2n (matches PRIMITIVE_ROOTS, b_field_element.rs:43-78; checked in tests)
static inline u64 hgl_root_of_unity(unsigned log2n) { return hgl_pow(7, (GL_P - 1) >> log2n); }
