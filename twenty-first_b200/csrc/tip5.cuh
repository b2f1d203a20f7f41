// tip5.cuh -- Tip5 permutation on raw Montgomery words, one hash per thread, state in registers.
//
// Replaces Tip5::permutation and its scalar/AVX-512 round functions
// (twenty-first/src/tip5/mod.rs:175-253, 529-533; tip5/avx512.rs:13-175).
// Per round (pinned by the reference KATs tip5/mod.rs:1145-1206,1294-1362):
//   lanes 0..3 : byte-wise LOOKUP_TABLE on the little-endian bytes of the raw word  (:197-207)
//   lanes 4..15: raw^7 mod p (valid on Montgomery words because R^6 = 1)            (:189-193)
//   t = circulant(MDS_MATRIX_FIRST_COLUMN) * s mod p                                (:154-157, naive.rs:54-68)
//   s = t + ROUND_CONSTANTS[16 r + i] * 2^64 mod p                                  (:68-149, 178-180)
// The MDS step follows the accumulator idea of the reference's AVX-512 path (avx512.rs:124-175):
// split every lane into 32-bit halves, accumulate the 16-bit x 32-bit products per half, seed the
// accumulators with the round-constant halves, reduce once.  On B200 the accumulation runs on the
// FP64 pipe: every partial sum is an integer below 2^52 (sum of the MDS column = 524757 < 2^20,
// times 2^32, plus a 32-bit round-constant half), so DFMA is exact, and it issues at one warp
// instruction per 2 cycles per sub-partition next to the integer pipes -- whereas the 512
// IMAD.WIDE per round of the integer formulation (one per 5.3 cycles, measured, tools/ubench.cu)
// made the whole permutation FMA-pipe bound.  Conversions use the 2^52 bias trick (no I2F/F2I).
#pragma once
#include "field.cuh"

#ifndef TIP5_CVT
#define TIP5_CVT 3  /* measured -2.4 % on the Merkle build (tools/ab.sh); bit 0: I2F instead of the 2^52 bias on input; bit 1: F2I on output */
#endif
#ifndef TIP5_MDS_CRT
#define TIP5_MDS_CRT 1  /* MDS as cyclic-8 + negacyclic-8 products (CRT over x^16 - 1): 320 instead of 512 FP64 ops per round */
#endif
#ifndef TIP5_ALU_FOLD
#define TIP5_ALU_FOLD 0  /* Solinas folds of the S-box products and the MDS outputs on the ALU pipe (field.cuh): measured +1 % time */
#endif
#ifndef TIP5_SQR
#define TIP5_SQR 1  /* the two squarings of x^7 with three wide multiplies (gl_sqr) */
#endif
#ifndef TIP5_MDS_SPLIT
#define TIP5_MDS_SPLIT 1
#endif
#ifndef TIP5_MUL_NW
#define TIP5_MUL_NW 0
#endif
#if TIP5_MUL_NW
#define TIP5_MUL gl_mul_nw
#define TIP5_REDUCE96 gl_reduce96
#elif TIP5_ALU_FOLD
#define TIP5_MUL gl_mul_alu
#define TIP5_REDUCE96 gl_reduce96a
#else
#define TIP5_MUL gl_mul
#define TIP5_REDUCE96 gl_reduce96
#endif
#define TIP5_STATE 16
#define TIP5_RATE 10
#define TIP5_ROUNDS 5
#define TIP5_DIGEST 5
#define TIP5_RAW_ONE 0xFFFFFFFFull /* BFieldElement::ONE as a raw word = 2^64 mod p */

// Filled by upload_tip5_constants (tip5_kernels.cuh): raw round constants split in 32-bit halves (with
// TIP5_MDS_CRT stored as the seeds of the two half-size products: [i] = (rc_i + rc_(i+8))/2,
// [8 + i] = (rc_i - rc_(i+8))/2, i < 8), and the S-box table.
__constant__ double c_tip5_rc_lo[TIP5_ROUNDS * TIP5_STATE];
__constant__ double c_tip5_rc_hi[TIP5_ROUNDS * TIP5_STATE];
// round-0 constants of fixed-length hashing: rc + sum_{j=10..15} M[(i-j)&15] * (ONE^7 mod p), per half
__constant__ double c_tip5_rc0f_lo[TIP5_STATE];
__constant__ double c_tip5_rc0f_hi[TIP5_STATE];
__constant__ uint8_t c_tip5_lut[256];
__constant__ u64 c_tip5_rc_raw[TIP5_ROUNDS * TIP5_STATE];  // raw round constants (cooperative kernels)

// MDS_MATRIX_FIRST_COLUMN, tip5/mod.rs:154-157
#define TIP5_MDS(k)                                                                              \
    ((k) == 0 ? 61402u : (k) == 1 ? 1108u : (k) == 2 ? 28750u : (k) == 3 ? 33823u : (k) == 4 ? 7454u \
     : (k) == 5 ? 43244u : (k) == 6 ? 53865u : (k) == 7 ? 12034u : (k) == 8 ? 56951u              \
     : (k) == 9 ? 27521u : (k) == 10 ? 41351u : (k) == 11 ? 40901u : (k) == 12 ? 12021u           \
     : (k) == 13 ? 59689u : (k) == 14 ? 26798u : 17845u)

#ifdef __CUDACC__

// copy the 256-byte S-box table into shared memory (call once per CTA, then __syncthreads)
// The table is computed, not copied: L[b] = ((b + 1)^3 + 256) mod 257 (tip5/mod.rs:1022-1053) is a handful of
// integer instructions, whereas `c_tip5_lut[threadIdx.x]` is a lane-divergent constant-bank read that the
// hardware serialises 32-fold (the kernel prologue showed up with 14 % of the stall samples in ncu).
__device__ __forceinline__ void tip5_load_lut(uint8_t *s_lut) {
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = (uint8_t)(((i + 1) * (i + 1) * (i + 1) + 256u) % 257u);
}

__device__ __forceinline__ u32 tip5_lut_word(u32 w, const uint8_t *s_lut) {
    u32 b0 = s_lut[w & 0xff];
    u32 b1 = s_lut[(w >> 8) & 0xff];
    u32 b2 = s_lut[(w >> 16) & 0xff];
    u32 b3 = s_lut[w >> 24];
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// One round.  NVAR = 16 in general; NVAR = 10 for the first round of fixed-length hashing
// (hash_10 / hash_pair, tip5/mod.rs:511-526, 559-586) where lanes 10..15 hold the constant ONE: their
// S-box outputs and MDS contributions are folded into the round-0 constants rc_lo / rc_hi passed in.
#ifndef TIP5_FUSED
#define TIP5_FUSED 0  /* MDS of round r and S-box of round r + 1 interleaved output pair by output pair (tip5_mds_sbox) */
#endif
#ifndef TIP5_FUSED_FENCE
#define TIP5_FUSED_FENCE 1
#endif
// 0.0 that ptxas cannot see through (a __device__ variable may be rewritten by the host): see tip5_mds_sbox
__device__ double g_tip5_zero = 0.0;

__device__ __forceinline__ u64 tip5_sbox_lut(u64 x, const uint8_t *s_lut) {
    return gl_pack(tip5_lut_word((u32)x, s_lut), tip5_lut_word((u32)(x >> 32), s_lut));
}
__device__ __forceinline__ u64 tip5_sbox_pow7(u64 x) {
#if TIP5_SQR
    const u64 x2 = gl_sqr(x);
    const u64 x4 = gl_sqr(x2);
#else
    const u64 x2 = TIP5_MUL(x, x);
    const u64 x4 = TIP5_MUL(x2, x2);
#endif
    return TIP5_MUL(x, TIP5_MUL(x2, x4));
}

// MDS of one round followed (SBOX) by the S-box of the NEXT round, output pair by output pair.
// In tip5_round below a warp runs an integer phase (S-box: ~1100 instructions bound by the quarter-rate wide
// multiplies of the FMA-heavy pipe) and then an FP64 phase (MDS); ncu showed the two pipes busy 63 % + 27 % of the
// time: the phases did not overlap, not even across the four warps of a scheduler.  Here the x^7 of lanes (i, i + 8)
// depends only on the 32 multiply-adds of output pair i, so the multiply-adds of the later pairs are independent
// work that fills the issue slots between the dependent integer instructions: every stretch of the instruction
// stream carries work for the FMA-heavy, the ALU and the FP64 pipe.  (Feeding the accumulators from the S-box side
// instead -- lane pair j into all 32 accumulators -- was tried first: ptxas sinks the multiply-adds, which then have
// no consumer until the end of the round, into one block behind the integer work.)
// in: s = S-box output of this round (NVAR variable lanes); out: s = S-box input of the next round, S-box applied
// when SBOX (lanes 0..3 canonical before their table look-up).
template <int NVAR, bool SBOX>
__device__ __forceinline__ void tip5_mds_sbox(u64 (&s)[TIP5_STATE], const uint8_t *s_lut, const double *rc_lo,
                                              const double *rc_hi, const double zero) {
    double al[8], ah[8], bl[8], bh[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double xl = __uint2double_rn((u32)s[j]), xh = __uint2double_rn((u32)(s[j] >> 32));
        if (j + 8 < NVAR) {
            const double yl = __uint2double_rn((u32)s[j + 8]), yh = __uint2double_rn((u32)(s[j + 8] >> 32));
            al[j] = xl + yl;
            bl[j] = xl - yl;
            ah[j] = xh + yh;
            bh[j] = xh - yh;
        } else {  // lane j + 8 is constant: folded into the seeds
            al[j] = bl[j] = xl;
            ah[j] = bh[j] = xh;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        // seeds: round-constant halves.  `zero * dep` (an opaque 0.0 times a finite double made of the S-box output
        // of pair i - 2) is a scheduling fence: without it ptxas issues all 256 multiply-adds of the round first (they
        // only depend on a / b) and the integer work afterwards -- the two phases this function exists to interleave.
        // With it the multiply-adds of pair i become ready when the S-box of pair i - 2 is done, i.e. they run next to
        // the S-box of pair i - 1 and are finished when that one ends.
        double pl = rc_lo[i], ph = rc_hi[i], ql = rc_lo[8 + i], qh = rc_hi[8 + i];
#if TIP5_FUSED_FENCE
        if (SBOX && i >= 2) {
            const double dep = __hiloint2double(0x43300000, (int)(u32)s[i - 2 + 8]);
            pl = fma(zero, dep, pl);
            ph = fma(zero, dep, ph);
            ql = fma(zero, dep, ql);
            qh = fma(zero, dep, qh);
        }
#endif
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = (i - j) & 7;
            const double mp = 0.5 * ((double)TIP5_MDS(k) + (double)TIP5_MDS(k + 8));
            const double mn = (j <= i ? 0.5 : -0.5) * ((double)TIP5_MDS(k) - (double)TIP5_MDS(k + 8));
            pl = fma(mp, al[j], pl);
            ph = fma(mp, ah[j], ph);
            ql = fma(mn, bl[j], ql);
            qh = fma(mn, bh[j], qh);
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const u64 acc_lo = __double2ull_rn(half ? pl - ql : pl + ql);
            const u64 acc_hi = __double2ull_rn(half ? ph - qh : ph + qh);
            const u64 x0 = acc_lo + (acc_hi << 32);
            const u32 x1 = (u32)(acc_hi >> 32) + (x0 < acc_lo ? 1u : 0u);
            const u64 v = TIP5_REDUCE96(x0, x1);
            const int lane = i + 8 * half;
            if (lane < 4) s[lane] = SBOX ? tip5_sbox_lut(gl_canon(v), s_lut) : gl_canon(v);
            else s[lane] = SBOX ? tip5_sbox_pow7(v) : v;
        }
    }
}

// NOUT: only lanes 0 .. NOUT-1 of the result are needed (the last round of hash_10 / hash_pair feeds the digest, lanes
// 0 .. 4, and nothing else: 5 of the 16 MDS outputs, tip5/mod.rs:583-585) -- the other outputs are left undefined.
template <int NVAR, int NOUT = TIP5_STATE>
__device__ __forceinline__ void tip5_round(u64 (&s)[TIP5_STATE], const uint8_t *s_lut, const double *rc_lo,
                                           const double *rc_hi) {
    // ---- S-box ----
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 lo = tip5_lut_word((u32)s[i], s_lut);
        u32 hi = tip5_lut_word((u32)(s[i] >> 32), s_lut);
        s[i] = gl_pack(lo, hi);
    }
#pragma unroll
    for (int i = 4; i < NVAR; i++) {
        u64 x = s[i];
#if TIP5_SQR
        u64 x2 = gl_sqr(x);
        u64 x4 = gl_sqr(x2);
#else
        u64 x2 = TIP5_MUL(x, x);
        u64 x4 = TIP5_MUL(x2, x2);
#endif
        u64 x6 = TIP5_MUL(x2, x4);
        s[i] = TIP5_MUL(x, x6);
    }
    // ---- MDS + round constants (exact integer arithmetic on the FP64 pipe) ----
    const double kBias = 4503599627370496.0;  // 2^52
#if TIP5_MDS_CRT && TIP5_MDS_SPLIT
    // Same CRT products as below, but the low 32-bit halves of all lanes first, then the high halves: 16 + 16
    // doubles live instead of 64, which is what lets the register allocation fit one more CTA per SM.
    u64 acc_l[TIP5_STATE];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        double a[8], b[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double x = __uint2double_rn(h ? (u32)(s[j] >> 32) : (u32)s[j]);
            if (j + 8 < NVAR) {
                const double y = __uint2double_rn(h ? (u32)(s[j + 8] >> 32) : (u32)s[j + 8]);
                a[j] = x + y;
                b[j] = x - y;
            } else {
                a[j] = b[j] = x;
            }
        }
        const double *rc = h ? rc_hi : rc_lo;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i >= NOUT) continue;  // neither lane i nor lane i + 8 is needed
            double p = rc[i], q = rc[8 + i];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = (i - j) & 7;
                const double mp = 0.5 * ((double)TIP5_MDS(k) + (double)TIP5_MDS(k + 8));
                const double mn = (j <= i ? 0.5 : -0.5) * ((double)TIP5_MDS(k) - (double)TIP5_MDS(k + 8));
                p = fma(mp, a[j], p);
                q = fma(mn, b[j], q);
            }
#pragma unroll
            for (int half = 0; half < 2; half++) {
                if (i + 8 * half >= NOUT) continue;
                const u64 acc = __double2ull_rn(half ? p - q : p + q);
                if (h == 0) {
                    acc_l[i + 8 * half] = acc;
                } else {
                    const u64 acc_lo = acc_l[i + 8 * half];
                    u64 x0 = acc_lo + (acc << 32);
                    u32 x1 = (u32)(acc >> 32) + (x0 < acc_lo ? 1u : 0u);
                    const u64 v = TIP5_REDUCE96(x0, x1);
                    s[i + 8 * half] = (i + 8 * half < 4) ? gl_canon(v) : v;
                }
            }
        }
    }
    return;
#endif
    double dl[NVAR], dh[NVAR];
#pragma unroll
    for (int j = 0; j < NVAR; j++) {
#if TIP5_CVT & 1
        dl[j] = __uint2double_rn((u32)s[j]);
        dh[j] = __uint2double_rn((u32)(s[j] >> 32));
#else
        dl[j] = __hiloint2double(0x43300000, (int)(u32)s[j]) - kBias;
        dh[j] = __hiloint2double(0x43300000, (int)(u32)(s[j] >> 32)) - kBias;
#endif
    }
#if TIP5_MDS_CRT
    // t(x) = M(x) s(x) mod (x^16 - 1) by the CRT split x^16 - 1 = (x^8 - 1)(x^8 + 1):
    //   a = s_lo + s_hi, b = s_lo - s_hi;  P = (M_lo + M_hi)/2 * a mod (x^8 - 1)  (cyclic),
    //   Q = (M_lo - M_hi)/2 * b mod (x^8 + 1) (negacyclic);  t_lo = P + Q, t_hi = P - Q.
    // 2 x 64 multiply-adds + 32 additions per half instead of 256 multiply-adds.  Exact: every partial sum is a
    // multiple of 1/2 with |2 P| <= t_i + t_(i+8) < 2^53 and |2 Q| < 2^53 (sum of the column = 524757 < 2^19.01,
    // inputs < 2^32, seeds = halves of sums/differences of the round constants).  rc_lo / rc_hi hold the seeds:
    // [0..8) for P = (rc_i + rc_(i+8))/2, [8..16) for Q = (rc_i - rc_(i+8))/2.
    double al[8], ah[8], bl[8], bh[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j + 8 < NVAR) {
            al[j] = dl[j] + dl[j + 8];
            bl[j] = dl[j] - dl[j + 8];
            ah[j] = dh[j] + dh[j + 8];
            bh[j] = dh[j] - dh[j + 8];
        } else {  // lane j + 8 is constant: folded into the seeds
            al[j] = bl[j] = dl[j];
            ah[j] = bh[j] = dh[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double pl = rc_lo[i], ph = rc_hi[i], ql = rc_lo[8 + i], qh = rc_hi[8 + i];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = (i - j) & 7;
            const double mp = 0.5 * ((double)TIP5_MDS(k) + (double)TIP5_MDS(k + 8));
            const double mn = (j <= i ? 0.5 : -0.5) * ((double)TIP5_MDS(k) - (double)TIP5_MDS(k + 8));
            pl = fma(mp, al[j], pl);
            ph = fma(mp, ah[j], ph);
            ql = fma(mn, bl[j], ql);
            qh = fma(mn, bh[j], qh);
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const double tl = half ? pl - ql : pl + ql;
            const double th = half ? ph - qh : ph + qh;
#if TIP5_CVT & 2
            const u64 acc_lo = __double2ull_rn(tl);
            const u64 acc_hi = __double2ull_rn(th);
#else
            const u64 acc_lo = (u64)__double_as_longlong(tl + kBias) & 0x000FFFFFFFFFFFFFull;
            const u64 acc_hi = (u64)__double_as_longlong(th + kBias) & 0x000FFFFFFFFFFFFFull;
#endif
            u64 x0 = acc_lo + (acc_hi << 32);
            u32 x1 = (u32)(acc_hi >> 32) + (x0 < acc_lo ? 1u : 0u);
            const u64 v = TIP5_REDUCE96(x0, x1);
            s[i + 8 * half] = (i + 8 * half < 4) ? gl_canon(v) : v;
        }
    }
#else
#pragma unroll
    for (int i = 0; i < TIP5_STATE; i++) {
        double al = rc_lo[i];
        double ah = rc_hi[i];
#pragma unroll
        for (int j = 0; j < NVAR; j++) {
            const double m = (double)TIP5_MDS((i - j) & 15);
            al = fma(m, dl[j], al);
            ah = fma(m, dh[j], ah);
        }
        // al, ah < 2^52: adding 2^52 leaves the integer in the 52 mantissa bits
#if TIP5_CVT & 2
        const u64 acc_lo = __double2ull_rn(al);
        const u64 acc_hi = __double2ull_rn(ah);
#else
        const u64 acc_lo = (u64)__double_as_longlong(al + kBias) & 0x000FFFFFFFFFFFFFull;
        const u64 acc_hi = (u64)__double_as_longlong(ah + kBias) & 0x000FFFFFFFFFFFFFull;
#endif
        // value = acc_lo + acc_hi * 2^32
        u64 x0 = acc_lo + (acc_hi << 32);
        u32 x1 = (u32)(acc_hi >> 32) + (x0 < acc_lo ? 1u : 0u);
        u64 v = gl_reduce96(x0, x1);
        s[i] = (i < 4) ? gl_canon(v) : v;
    }
#endif
}

// s: 16 raw words as stored by the caller (the S-box LUT acts on the raw bytes as they are, like
// split_and_lookup tip5/mod.rs:197-207; lanes 4..15 may be any representative mod p).
// FIXED: lanes 10..15 are known to hold raw ONE (their contents are ignored).
// On exit lanes 0..3 are canonical, lanes 4..15 weak; callers canonicalise what they store.
#ifndef TIP5_DIGEST_LAST
#define TIP5_DIGEST_LAST 0  /* last round of the fixed-length hashes peeled, with only the five digest lanes of its MDS (-250 instructions of 7760 per hash): measured SLOWER, Merkle 2^24 5.57 against 5.34 ms, hash_10 3.07 against 3.19 G/s (profiles/r02k_ab_tip5_digest_only_last_round.txt) -- a third copy of the round in the instruction stream and 120 instead of 109 registers cost more than the 3 % of work saved */
#endif
template <bool FIXED = false, bool DIGEST_ONLY = false>
__device__ __forceinline__ void tip5_permutation(u64 (&s)[TIP5_STATE], const uint8_t *s_lut) {
#if TIP5_FUSED && TIP5_MDS_CRT
    // S-box(0) | MDS(0) + S-box(1) | ... | MDS(3) + S-box(4) | MDS(4)
#pragma unroll
    for (int i = 0; i < (FIXED ? TIP5_RATE : TIP5_STATE); i++) s[i] = i < 4 ? tip5_sbox_lut(s[i], s_lut) : tip5_sbox_pow7(s[i]);
    const double zero = *(const volatile double *)&g_tip5_zero;
    if (FIXED) tip5_mds_sbox<TIP5_RATE, true>(s, s_lut, c_tip5_rc0f_lo, c_tip5_rc0f_hi, zero);
#pragma unroll 1
    for (int r = FIXED ? 1 : 0; r < TIP5_ROUNDS - 1; r++)
        tip5_mds_sbox<TIP5_STATE, true>(s, s_lut, c_tip5_rc_lo + r * TIP5_STATE, c_tip5_rc_hi + r * TIP5_STATE, zero);
    tip5_mds_sbox<TIP5_STATE, false>(s, s_lut, c_tip5_rc_lo + (TIP5_ROUNDS - 1) * TIP5_STATE,
                                     c_tip5_rc_hi + (TIP5_ROUNDS - 1) * TIP5_STATE, zero);
    return;
#endif
    if (FIXED) tip5_round<TIP5_RATE>(s, s_lut, c_tip5_rc0f_lo, c_tip5_rc0f_hi);
    constexpr int kLoopEnd = (DIGEST_ONLY && TIP5_DIGEST_LAST && TIP5_MDS_CRT && TIP5_MDS_SPLIT) ? TIP5_ROUNDS - 1 : TIP5_ROUNDS;
#pragma unroll 1
    for (int r = FIXED ? 1 : 0; r < kLoopEnd; r++)
        tip5_round<TIP5_STATE>(s, s_lut, c_tip5_rc_lo + r * TIP5_STATE, c_tip5_rc_hi + r * TIP5_STATE);
    if (kLoopEnd < TIP5_ROUNDS)
        tip5_round<TIP5_STATE, TIP5_DIGEST>(s, s_lut, c_tip5_rc_lo + (TIP5_ROUNDS - 1) * TIP5_STATE,
                                            c_tip5_rc_hi + (TIP5_ROUNDS - 1) * TIP5_STATE);
}

// ---- cooperative form: 16 lanes per hash -----------------------------------------------------------------
// For the latency-bound top of a Merkle tree (a level with few nodes is one dependent ~10 us single-thread hash
// after the other): lane i of a 16-lane group holds state element i, the S-box of a round is one x^7 (or one
// LUT word) per lane, and the MDS row of lane i is gathered by 15 rotating shuffles,
// t_i = sum_k M[k] s_((i - k) & 15), so that the matrix entry is the same immediate for all lanes.  About 4x
// lower latency per hash than the one-thread form and ~3x less throughput: only used below kMerkleCoopCnt.
// rc_raw: the 80 raw round constants (shared memory).  Returns the new state element of this lane, canonical.
__device__ __forceinline__ u64 tip5_permutation_coop(u64 s, u32 lane16, const uint8_t *s_lut, const u64 *rc_raw) {
    // the two 16-lane groups of a warp synchronise separately (one of them may have left the kernel)
    const u32 mask = 0xffffu << (threadIdx.x & 16);
#pragma unroll 1
    for (int r = 0; r < TIP5_ROUNDS; r++) {
        if (lane16 < 4) {
            const u32 lo = tip5_lut_word((u32)s, s_lut);
            const u32 hi = tip5_lut_word((u32)(s >> 32), s_lut);
            s = gl_pack(lo, hi);
        } else {
            const u64 x2 = gl_sqr(s);
            const u64 x4 = gl_sqr(x2);
            s = gl_mul(s, gl_mul(x2, x4));
        }
        const u32 s_lo = (u32)s, s_hi = (u32)(s >> 32);
        u64 acc_lo = (u64)s_lo * TIP5_MDS(0), acc_hi = (u64)s_hi * TIP5_MDS(0);
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const u32 o_lo = __shfl_sync(mask, s_lo, (lane16 - k) & 15, 16);
            const u32 o_hi = __shfl_sync(mask, s_hi, (lane16 - k) & 15, 16);
            acc_lo += (u64)o_lo * TIP5_MDS(k);  // sums of sixteen 48-bit products: < 2^52
            acc_hi += (u64)o_hi * TIP5_MDS(k);
        }
        const u64 x0 = acc_lo + (acc_hi << 32);
        const u32 x1 = (u32)(acc_hi >> 32) + (x0 < acc_lo ? 1u : 0u);
        s = gl_add(gl_canon(gl_reduce96(x0, x1)), rc_raw[r * TIP5_STATE + lane16]);
    }
    return s;
}

#endif
