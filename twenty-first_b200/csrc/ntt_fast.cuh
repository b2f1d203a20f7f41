// ntt_fast.cuh -- 1024-point passes held in registers: the hot kernels of the batched 2^20 NTT.
//
// Same pass structure and index algebra as ntt_kernels.cuh (column pass / transposing row pass),
// specialised to N_p = 1024 = 32 x 32:
//
//   * one warp owns one 1024-point column; every lane keeps 32 elements in registers;
//   * step A: 32-point DFT over the high index digit, step B: 32-point DFT over the low digit; all
//     butterflies inside a 32-point DFT use power-of-two twiddles (omega_32 = 2^78 mod p, because
//     omega_64 = 2^39, b_field_element.rs:43-78) -> shifts + Solinas folds, no 64x64 multiply;
//   * the 32-point DFT is decimation in time on lazy values: the twiddled operand is made canonical
//     (gl_shlc / gl_canonw), sums and differences stay "any u64" (gl_addl / gl_subl: two adds + a
//     predicated two-add wrap correction), so a butterfly is 8 instructions + the twiddle;
//   * one general multiply per element between the two steps (omega_1024^(b k1)) and, for column
//     passes, one for the inter-pass twiddle omega_B^(i j_rest), read from a full [j_rest][i] table in
//     L2 (coalesced, 8 MiB for B = 2^20).  Both steps run through the same loop body (halves the
//     instruction footprint: the unrolled 32-point DFT is ~1400 instructions);
//   * the 32 x 32 transpose between the steps goes through the warp's own slice of shared memory
//     (__syncwarp only); __syncthreads is needed only around the coalesced global staging;
//   * the batched 2^20 transform takes template specialisations without index arithmetic or uniform
//     branches in the staging loops (PLAIN column pass with cp.async staging, POST-less row pass),
//     grids without index divisions and 32-bit row offsets.
//
// Values in flight are lazy (any u64 representative); only the last pass canonicalises on store, so
// non-canonical input words are accepted as well.
//
// Shared-memory layout: tile[col][S] u64 with S = 1092 for 4 columns per CTA (>= 32 x 34 for the transpose with
// 16-byte aligned rows; = 4 mod 16: the 16 lanes of a
// half warp -- 4 columns x 4 rows of the staging pattern -- hit 16 different 8-byte bank pairs; the per-warp
// column patterns are unit-stride or stride 33).
#pragma once
#include "ntt_kernels.cuh"
#include "tma.cuh"

namespace tf21 {

#ifndef TF21_SHL_SINGLE
#define TF21_SHL_SINGLE 1  /* shift form of the single-pass 2^10 kernel (no staging, lighter ALU load): the wide-multiply form is 4 % faster there */
#endif
#ifndef TF21_ROW_STORE128
#define TF21_ROW_STORE128 1  /* 16-byte stores of column pairs in the stage-out loops: 2.92 -> 2.90 ms */
#endif
#ifndef TF21_TRANSPOSE_LDS128
#define TF21_TRANSPOSE_LDS128 1  /* 32 x 32 transpose with row stride 34 and 16-byte read-back: 2.90 -> 2.85 ms */
#endif
#ifndef TF21_COL_STORE128
#define TF21_COL_STORE128 0
#endif
#ifndef TF21_COL_MASKMUL
#define TF21_COL_MASKMUL false  /* true: mask-style wrap corrections in the products of the column pass (it spilled with the predicated form before the PLAIN specialisation; now predicated is 2 % faster: 3.02 against 3.09 ms) */
#endif
#ifndef TF21_SHL_COL
#define TF21_SHL_COL TF21_SHL_WIDE
#endif
#ifndef TF21_SHL_ROW
#define TF21_SHL_ROW TF21_SHL_WIDE
#endif
#ifndef TF21_FIXW_COL
#define TF21_FIXW_COL 0
#endif
#ifndef TF21_FIXW_ROW
#define TF21_FIXW_ROW 0
#endif
#ifndef TF21_TW_PREFETCH
#define TF21_TW_PREFETCH 0  /* measured: 1.356 against 1.341 ms for the column pass (L1 too small next to 4 x 36 KB of shared memory) */
#endif
#ifndef TF21_DFT_EXIT_MID
#define TF21_DFT_EXIT_MID 0
#endif
// A/B knobs for the resident-CTA target of the thread-per-column passes (profiles/r02k_ab_occupancy_small_col_pruned.txt: no
// gain); unset = plain __launch_bounds__(128) -- a minimum of ONE block is not the same thing to ptxas: it then spends up
// to 150 registers on these kernels (2^14: 0.98 -> 1.005 ms per GiB, pruned pass 0.50 -> 0.53 ms)
#ifdef TF21_SMALLCOL_BLOCKS
#define TF21_SMALLCOL_BOUNDS __launch_bounds__(128, TF21_SMALLCOL_BLOCKS)
#else
#define TF21_SMALLCOL_BOUNDS __launch_bounds__(128)
#endif
#ifdef TF21_PRUNED_BLOCKS
#define TF21_PRUNED_BOUNDS __launch_bounds__(128, TF21_PRUNED_BLOCKS)
#else
#define TF21_PRUNED_BOUNDS __launch_bounds__(128)
#endif
#ifndef TF21_MID_DEFAULT_MASK
#define TF21_MID_DEFAULT_MASK 0x3e0  /* K = 5 .. 9: measured faster than the thread-per-column passes they replace (ms per GiB of 2^15 / 16 / 17 / 18 / 19 / 26 / 27: 1.12 / 1.36 / 1.39 / 1.42 / 1.63 / 2.17 / 2.28 -> 1.06 / 1.09 / 1.21 / 1.28 / 1.38 / 1.90 / 2.08); K = 9 is the two-thread form ntt_mid9_col_kernel (the generic form with 32 elements per thread is slower than the old passes: 1.64) */
#endif
#ifndef TF21_FAST_COLS
#define TF21_FAST_COLS 4  /* 4 CTAs of 128 threads per SM: finer interleaving of staging and compute phases than 2 x 256 (tools/ab.sh: 3.14 ms against 3.27 ms per 256-column batch once the staging is asynchronous; 32-byte row segments = one DRAM sector) */
#endif
constexpr u32 kFastCols = TF21_FAST_COLS;  // word-columns (= warps) per CTA: 4, 8 or 16
// u64 per column slice in shared memory (>= 32 * 33), chosen so that the staging pattern
// (kFastCols lanes per row segment) is bank-conflict free: S mod 16 = 16 / kFastCols
#if TF21_TRANSPOSE_LDS128
constexpr u32 kTransposeStride = 34;
constexpr u32 kFastS = kFastCols == 4 ? 1092 : kFastCols == 8 ? 1090 : 1089;
#else
constexpr u32 kTransposeStride = 33;
constexpr u32 kFastS = kFastCols == 4 ? 1060 : kFastCols == 8 ? 1058 : 1057;
#endif
constexpr u32 kFastThreads = kFastCols * 32;
#ifndef TF21_MIN_BLOCKS
#define TF21_MIN_BLOCKS (16 / TF21_FAST_COLS)
#endif
constexpr u32 kFastMinBlocks = TF21_MIN_BLOCKS;  // 512 threads of 128 registers per SM (the 32 x u64 column of a lane needs 64)
constexpr u32 kStageRowsPerIt = 32 / kFastCols;  // rows covered by one warp instruction of the staging loops
constexpr u32 kStageRowsPerWarp = 1024 / kFastCols;
constexpr size_t kFastSmem = (size_t)kFastCols * kFastS * sizeof(u64);

__host__ __device__ constexpr u32 brev5(u32 k) {
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}
__host__ __device__ constexpr u32 brev_bits(u32 k, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; i++) r |= ((k >> i) & 1u) << (bits - 1 - i);
    return r;
}

// 2^A-point DFT in registers (A <= 6), decimation in time: in v[a] = x[a] (any u64), out
// v[brev_A(k)] = sum_a x[a] w^(a k) (any u64), w = omega_{2^A}^(+-1) = 2^(+-39 * 2^(6-A)) because
// omega_64 = 2^39 (PRIMITIVE_ROOTS, b_field_element.rs:43-78).  (butterfly of ntt.rs:203-210; the
// twiddles are shifts; exponents >= 96 use 2^96 = -1 and swap the roles of sum and difference)
// One butterfly per template instance (LS = stage, IDX = butterfly index inside the stage): every
// register index and every shift amount is a compile-time constant by construction, so the value
// array can never fall back to local memory (a `#pragma unroll` loop nest around inline asm with
// labels was left partially rolled by the compiler for some sizes).
template <bool INV, int A, int LS, int IDX, int SHLV>
__device__ __forceinline__ void dft_pow2_step(u64 (&v)[1 << A], u32 one) {
    constexpr int N = 1 << A;
    if constexpr (LS > A) {
        return;
    } else if constexpr (IDX >= N / 2) {
        dft_pow2_step<INV, A, LS + 1, 0, SHLV>(v, one);
    } else {
        constexpr int EU = (39 << (6 - A)) % 192;
        constexpr int m = 1 << LS, half = m >> 1;
        constexpr int k = (IDX / half) * m, j = IDX % half;
        constexpr int iu = brev_bits(k + j, A), ib = brev_bits(k + j + half, A);
        constexpr int E0 = (EU * j * (N / m)) % 192;
        constexpr int E = INV ? (192 - E0) % 192 : E0;
        constexpr bool neg = E >= 96;
        constexpr int S = neg ? E - 96 : E;
        u64 t;
        if constexpr (S == 0) t = (TF21_CANON_WIDE & 1) ? gl_canon_wide(v[ib], one) : gl_canonw(v[ib]);
        else t = gl_shlc<S, SHLV>(v[ib], one);
        const u64 u = v[iu];
        if constexpr (!neg) {
            v[iu] = TF21_ADDFIX_WIDE ? gl_addp_wide(u, t, one) : gl_addl(u, t);
            v[ib] = gl_subl(u, t, one);
        } else {
            v[iu] = gl_subl(u, t, one);
            v[ib] = TF21_ADDFIX_WIDE ? gl_addp_wide(u, t, one) : gl_addl(u, t);
        }
        dft_pow2_step<INV, A, LS, IDX + 1, SHLV>(v, one);
    }
}

template <bool INV, int A, int SHLV = TF21_SHL_WIDE>
__device__ __forceinline__ void dft_pow2(u64 (&v)[1 << A], u32 one = c_gl_one) {
    dft_pow2_step<INV, A, 1, 0, SHLV>(v, one);
}

// ---- optimistic canonicalisation -------------------------------------------------------------------------------
// The lazy butterflies need their twiddled operand t <= p.  A lazy value (a sum, a difference, a folded product or
// shift) is >= p only when its high word is 0xffffffff -- one value in 2^32 for random data -- yet dft_pow2_step pays
// a canonicalisation (2 ALU + 2 FMA-pipe instructions) for each of the 31 operands with a trivial twiddle and inside
// each of the 13 shifts by less than 32 bits of a 32-point transform.  Here a stage first forms all its twiddled
// operands in place (shifts by < 32 bits only folded), takes the maximum of the high words that are not known to be
// canonical (one three-input VIMNMX per two operands), and only a warp in which some lane saw 0xffffffff
// canonicalises them (a warp-uniform branch around the old code); then the 16 add / subtract pairs follow.
// 176 -> ~40 instructions per 32-point transform; bit-exact for every input (the slow path is the old path; adversarial
// words in tests/test_gpu_parity.py take it).
#ifndef TF21_OPT_CANON
#define TF21_OPT_CANON 1
#endif
template <bool INV, int A, int LS, int IDX>
struct DftBfly {
    static constexpr int N = 1 << A;
    static constexpr int EU = (39 << (6 - A)) % 192;
    static constexpr int m = 1 << LS, half = m >> 1;
    static constexpr int k = (IDX / half) * m, j = IDX % half;
    static constexpr int iu = brev_bits(k + j, A), ib = brev_bits(k + j + half, A);
    static constexpr int E0 = (EU * j * (N / m)) % 192;
    static constexpr int E = INV ? (192 - E0) % 192 : E0;
    static constexpr bool neg = E >= 96;
    static constexpr int S = neg ? E - 96 : E;
    static constexpr bool flagged = S < 32;  // trivial twiddle or a shift that is only folded
};
// Running flag over the high words of a stage: "some word was 0xffffffff".
// TF21_OPT_FLAG_MUL: the flag is the product of the (hi + 1) mod 2^32 -- zero as soon as one factor is zero -- formed
// by one IMAD per word on the FMA-heavy pipe (the kernels are ALU-pipe bound) instead of a VIMNMX per one or two words
// on the ALU pipe.  A product can also vanish because its factors bring 32 factors of two together (for random words
// about once in several hundred stages): that only sends a warp through the slow path, which is always correct.
#ifndef TF21_OPT_FLAG_MUL
#define TF21_OPT_FLAG_MUL 0  /* measured neutral (2.319 against 2.324 ms) with chains of <= 8 factors, +2.5 % with chains of 16 (false hits in one of 32 lanes) */
#endif
#if TF21_OPT_FLAG_MUL
constexpr u32 kOptFlagInit = 1u;
__device__ __forceinline__ u32 dft_opt_flag(u32 acc, u32 hi) {
    u32 r;
    asm("mad.lo.u32 %0,%1,%2,%1;" : "=r"(r) : "r"(acc), "r"(hi));  // acc * (hi + 1)
    return r;
}
__device__ __forceinline__ bool dft_opt_hit(u32 acc) { return acc == 0u; }
#else
constexpr u32 kOptFlagInit = 0u;
__device__ __forceinline__ u32 dft_opt_flag(u32 acc, u32 hi) { return max(acc, hi); }
__device__ __forceinline__ bool dft_opt_hit(u32 acc) { return acc == 0xffffffffu; }
#endif
// mx[2]: the first eight and the second eight butterflies of a stage keep separate flags (a product of more than
// eight factors collects 32 factors of two in one of the 32 lanes of a warp far too often)
template <bool INV, int A, int LS, int IDX, int SHLV>
__device__ __forceinline__ void dft_opt_twiddle(u64 (&v)[1 << A], u32 one, u32 (&mxs)[2]) {
    if constexpr (IDX < (1 << A) / 2) {
        using B = DftBfly<INV, A, LS, IDX>;
        u32 &mx = mxs[TF21_OPT_FLAG_MUL ? IDX / 8 : 0];
        if constexpr (B::S != 0) v[B::ib] = gl_shlc<(B::S ? B::S : 1), (SHLV & 15), true>(v[B::ib], one);
        if constexpr (B::flagged) mx = dft_opt_flag(mx, (u32)(v[B::ib] >> 32));
        dft_opt_twiddle<INV, A, LS, IDX + 1, SHLV>(v, one, mxs);
    }
}
template <bool INV, int A, int LS, int IDX>
__device__ __forceinline__ void dft_opt_fix(u64 (&v)[1 << A]) {
    if constexpr (IDX < (1 << A) / 2) {
        using B = DftBfly<INV, A, LS, IDX>;
        if constexpr (B::flagged) v[B::ib] = gl_canonw(v[B::ib]);
        dft_opt_fix<INV, A, LS, IDX + 1>(v);
    }
}
// FIXW = K > 0: every K-th sum takes its wrap correction as one predicated IMAD.WIDE.U32 (one * EPS + s, FMA-heavy
// pipe, 5.3 cycles) instead of IADD3 + IMAD.X (2 cycles of the ALU pipe + 2 of the FMA-heavy pipe)
template <bool INV, int A, int LS, int IDX, int FIXW>
__device__ __forceinline__ void dft_opt_addsub(u64 (&v)[1 << A], u32 one) {
    if constexpr (IDX < (1 << A) / 2) {
        using B = DftBfly<INV, A, LS, IDX>;
        const u64 u = v[B::iu], t = v[B::ib];
        constexpr bool wide = FIXW > 0 && (IDX % (FIXW > 0 ? FIXW : 1)) == 0;
        if constexpr (!B::neg) {
            v[B::iu] = wide ? gl_addp_wide(u, t, one) : gl_addl(u, t);
            v[B::ib] = gl_subl(u, t, one);
        } else {
            v[B::iu] = gl_subl(u, t, one);
            v[B::ib] = wide ? gl_addp_wide(u, t, one) : gl_addl(u, t);
        }
        dft_opt_addsub<INV, A, LS, IDX + 1, FIXW>(v, one);
    }
}
// must be called by all 32 lanes of a warp together
template <bool INV, int A, int LS, int SHLV>
__device__ __forceinline__ void dft_opt_stage(u64 (&v)[1 << A], u32 one) {
    if constexpr (LS <= A) {
        u32 mx[2] = {kOptFlagInit, kOptFlagInit};
        dft_opt_twiddle<INV, A, LS, 0, SHLV>(v, one, mx);
        if (__builtin_expect(__any_sync(0xffffffffu, dft_opt_hit(mx[0]) || (TF21_OPT_FLAG_MUL && dft_opt_hit(mx[1]))), 0))
            dft_opt_fix<INV, A, LS, 0>(v);
        dft_opt_addsub<INV, A, LS, 0, (SHLV >> 4)>(v, one);
        dft_opt_stage<INV, A, LS + 1, SHLV>(v, one);
    }
}

template <bool INV, int SHLV = TF21_SHL_WIDE>
__device__ __forceinline__ void dft32(u64 (&v)[32], u32 one = c_gl_one) {
    dft_pow2<INV, 5, SHLV>(v, one);
}
// the 32-point transform of the warp-per-column passes (whole warps only)
template <bool INV, int SHLV = TF21_SHL_WIDE>
__device__ __forceinline__ void dft32_warp(u64 (&v)[32], u32 one = c_gl_one) {
#if TF21_OPT_CANON
    dft_opt_stage<INV, 5, 1, SHLV>(v, one);  // SHLV: shift form + 16 * FIXW
#else
    dft_pow2<INV, 5, (SHLV & 15)>(v, one);
#endif
}

// the same butterflies on the 2^A consecutive registers v[OFF .. OFF + 2^A) of a 32-register column
template <bool INV, int A, int OFF, int LS, int IDX, int SHLV, int STRIDE = 1>
__device__ __forceinline__ void dft_sub_step(u64 (&v)[32], u32 one) {
    constexpr int N = 1 << A;
    if constexpr (LS > A) {
        return;
    } else if constexpr (IDX >= N / 2) {
        dft_sub_step<INV, A, OFF, LS + 1, 0, SHLV, STRIDE>(v, one);
    } else {
        constexpr int EU = (39 << (6 - A)) % 192;
        constexpr int m = 1 << LS, half = m >> 1;
        constexpr int k = (IDX / half) * m, j = IDX % half;
        constexpr int iu = OFF + STRIDE * brev_bits(k + j, A), ib = OFF + STRIDE * brev_bits(k + j + half, A);
        constexpr int E0 = (EU * j * (N / m)) % 192;
        constexpr int E = INV ? (192 - E0) % 192 : E0;
        constexpr bool neg = E >= 96;
        constexpr int S = neg ? E - 96 : E;
        u64 t;
        if constexpr (S == 0) t = (TF21_CANON_WIDE & 1) ? gl_canon_wide(v[ib], one) : gl_canonw(v[ib]);
        else t = gl_shlc<S, SHLV>(v[ib], one);
        const u64 u = v[iu];
        if constexpr (!neg) {
            v[iu] = TF21_ADDFIX_WIDE ? gl_addp_wide(u, t, one) : gl_addl(u, t);
            v[ib] = gl_subl(u, t, one);
        } else {
            v[iu] = gl_subl(u, t, one);
            v[ib] = TF21_ADDFIX_WIDE ? gl_addp_wide(u, t, one) : gl_addl(u, t);
        }
        dft_sub_step<INV, A, OFF, LS, IDX + 1, SHLV, STRIDE>(v, one);
    }
}
// 32 / 2^A independent 2^A-point DFTs on the STRIDED register groups v[o + s 2^(5-A)], s < 2^A, o < 2^(5-A)
// (the DFT runs over the HIGH bits of the register index); out: v[o + brev_A(k) 2^(5-A)]
template <bool INV, int A, int O = 0>
__device__ __forceinline__ void dft_groups_strided(u64 (&v)[32], u32 one) {
    if constexpr (A > 0 && O < (32 >> A)) {
        dft_sub_step<INV, A, O, 1, 0, TF21_SHL_WIDE, (32 >> A)>(v, one);
        dft_groups_strided<INV, A, O + 1>(v, one);
    }
}
// 32 / 2^A independent 2^A-point DFTs on the register groups v[g 2^A .. (g + 1) 2^A); out: v[g 2^A + brev_A(k)]
template <bool INV, int A, int G = 0>
__device__ __forceinline__ void dft_groups(u64 (&v)[32], u32 one) {
    if constexpr (A > 0 && G < (32 >> A)) {
        dft_sub_step<INV, A, G << A, 1, 0, TF21_SHL_WIDE>(v, one);
        dft_groups<INV, A, G + 1>(v, one);
    }
}

// lazy twiddle from split tables (any u64 representative)
__device__ __forceinline__ u64 scale_factor_l(const ScaleTab &t, u64 idx) {
    return gl_mul(__ldg(t.lo + (idx & ((1ull << t.h) - 1))), __ldg(t.hi + (idx >> t.h)));
}

// The 1024-point DFT of the column held by this warp.
// in : v[a] = x[32 a + lane]                       (natural order, any u64)
// out: slice[i] = X[i] * (tw1 ? tw1[32 (i >> 5)] : 1), i = lane + 32 k2   (any u64)
// slice: this warp's kFastS-word shared-memory slice; tw0 = t1 + lane with t1[k1*32+b] = w1024^(k1 b);
// tw1: per-lane pointer into the inter-pass twiddle row (or nullptr).
// TILE_OUT (TMA column pass): the second step stores X[lane + 32 k2] at out1[k2 * ss1] -- the swizzled tile that
// goes back to HBM as one box store -- after a CTA barrier, because that tile overlaps the private transpose
// slices of the other warps.
template <bool INV, bool MASKMUL = false, int SHLV = TF21_SHL_WIDE, bool TILE_OUT = false, bool CANON_OUT = false>
__device__ __forceinline__ void dft1024_warp(u64 (&v)[32], u64 *slice, const u64 *tw0, const u64 *tw1, u32 lane,
                                             u64 *out1 = nullptr, u32 ss1 = 32u) {
    // 1 in a register of its own: t1[0] = omega_1024^0 (see gl_subp); a global load is never re-issued by ptxas
    const u32 one = (u32)__ldg(tw0 - lane);
#if TF21_DFT_EXIT_MID
    // Same two steps with the loop left in the middle: the output phase of the second step sits behind the loop,
    // where the compiler knows that v[] is dead (in the rolled form below every canonicalised / multiplied word is
    // first copied, because v[] looks live across the back edge: 8 instead of 4 instructions per stored word)
#pragma unroll 1
    for (int it = 0;; it++) {
        dft32_warp<INV, SHLV>(v, one);
        if (it) break;
        u64 *out = slice + lane;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if (k == 0) {  // t1[0][b] = omega_1024^0 = 1
                out[0] = v[0];
                continue;
            }
            out[k * kTransposeStride] =
                MASKMUL ? gl_mul_mask(v[brev5(k)], __ldg(tw0 + 32 * k)) : gl_mul(v[brev5(k)], __ldg(tw0 + 32 * k));
        }
        __syncwarp();
        const ulonglong2 *rowp = reinterpret_cast<const ulonglong2 *>(slice + lane * kTransposeStride);
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const ulonglong2 p = rowp[b];
            v[2 * b] = p.x;
            v[2 * b + 1] = p.y;
        }
    }
    {
        u64 *out = TILE_OUT ? out1 : slice + lane;
        const u32 ss = TILE_OUT ? ss1 : 32u;
        if (TILE_OUT) __syncthreads();  // every warp has read its transposed column back
        __syncwarp();
        if (tw1) {
#pragma unroll
            for (int k = 0; k < 32; k++)
                out[k * ss] = MASKMUL ? gl_mul_mask(v[brev5(k)], __ldg(tw1 + 32 * k)) : gl_mul(v[brev5(k)], __ldg(tw1 + 32 * k));
        } else if (CANON_OUT) {  // last pass: canonical words straight into the outgoing tile
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = gl_canonw(v[brev5(k)]);
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = v[brev5(k)];
        }
        __syncwarp();
    }
    return;
#endif
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        dft32_warp<INV, SHLV>(v, one);
        const u64 *tw = it ? tw1 : tw0;
        u64 *out = (TILE_OUT && it) ? out1 : slice + lane;
        const u32 ss = it ? (TILE_OUT ? ss1 : 32u) : kTransposeStride;
        if (TILE_OUT && it) __syncthreads();  // every warp has read its transposed column back
        __syncwarp();
        if (tw) {
#pragma unroll
            for (int k = 0; k < 32; k++) {
                if (k == 0 && it == 0) {  // t1[0][b] = omega_1024^0 = 1
                    out[0] = v[0];
                    continue;
                }
                out[k * ss] = MASKMUL ? gl_mul_mask(v[brev5(k)], __ldg(tw + 32 * k)) : gl_mul(v[brev5(k)], __ldg(tw + 32 * k));
            }
        } else if (CANON_OUT && it) {  // last pass: canonical words straight into the outgoing tile
#if TF21_OPT_CANON
            // optimistic: one maximum over the 32 high words; the canonicalisation only runs in a warp that saw 0xffffffff
            // (two chains: 32 factors of a single product would bring 32 factors of two together too often)
            u32 mx[4] = {kOptFlagInit, kOptFlagInit, kOptFlagInit, kOptFlagInit};
#pragma unroll
            for (int k = 0; k < 32; k++) mx[TF21_OPT_FLAG_MUL ? k & 3 : 0] = dft_opt_flag(mx[TF21_OPT_FLAG_MUL ? k & 3 : 0], (u32)(v[k] >> 32));
            if (__builtin_expect(__any_sync(0xffffffffu, dft_opt_hit(mx[0]) || (TF21_OPT_FLAG_MUL && (dft_opt_hit(mx[1]) || dft_opt_hit(mx[2]) || dft_opt_hit(mx[3])))), 0)) {
#pragma unroll
                for (int k = 0; k < 32; k++) v[k] = gl_canonw(v[k]);
            }
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = v[brev5(k)];
#else
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = (TF21_CANON_WIDE & 2) ? gl_canon_wide(v[brev5(k)], one) : gl_canonw(v[brev5(k)]);
#endif
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = v[brev5(k)];
        }
        __syncwarp();
        if (it == 0) {
#if TF21_TRANSPOSE_LDS128
            // row stride 34: every row starts 16-byte aligned, the read-back takes 16 LDS.128 instead of 32 LDS.64
            const ulonglong2 *rowp = reinterpret_cast<const ulonglong2 *>(slice + lane * kTransposeStride);
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const ulonglong2 p = rowp[b];
                v[2 * b] = p.x;
                v[2 * b + 1] = p.y;
            }
#else
#pragma unroll
            for (int b = 0; b < 32; b++) v[b] = slice[lane * kTransposeStride + b];
#endif
        }
    }
}

struct FastColArgs {
    const u64 *src;
    u64 *dst;
    u64 src_array_words, dst_array_words;
    u32 w;
    u64 inner_words;
    u32 n_col_tiles, n_outer;
    u64 n_in_elems;
    const u64 *t1;     // [32][32] omega_1024^(k1 b)
    const u64 *tw_full;  // [inner_elems][1024] omega_B^(j i) (times n^-1 for inverse pass 1) or nullptr
    ScaleTab tw;       // split tables, used when tw_full == nullptr
    u32 log_b;
    u64 tw_scalar;     // extra factor folded into the split path (n^-1), 0 => none
    ScaleTab pre;
};

// PLAIN: every input row exists (n_in == n), no coset pre-scale and the inter-pass twiddles come from the full
// table -- the batched 2^20 transform.  The staging loops then carry no index arithmetic or uniform branches
// (ncu: ~890 of 5300 instructions per warp and most `no_instruction` stalls of the general form were staging).
template <bool INV, bool PLAIN>
__global__ void __launch_bounds__(kFastThreads, kFastMinBlocks) ntt1024_col_kernel(const FastColArgs a) {
    extern __shared__ __align__(16) u64 smem[];
    u64 *tile = smem;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // grid = (column tiles, outer blocks, arrays of the batch): no index divisions
    const u32 ct = blockIdx.x, o = blockIdx.y, b = blockIdx.z;
    const u64 q0 = (u64)ct * kFastCols;
    const u64 inner_elems = a.w == 1 ? (u32)a.inner_words : (u32)a.inner_words / 3u;
    const u32 c = lane % kFastCols, rsub = lane / kFastCols;
    // a tile spans 1024 * inner_words words; inner_words = 1024 w in every plan (the column pass always works on
    // 2^20-point blocks), so row offsets fit 32 bits: one narrow IMAD per row instead of a 64-bit multiply
    const u32 row_words = (u32)a.inner_words;

    // ---- stage in: kFastCols lanes per row segment (32 bytes for 4 columns), 32 / kFastCols rows per warp instruction ----
    const u64 block_off = (u64)o * 1024 * a.inner_words + q0;
    {
        const u64 *src = a.src + (u64)b * a.src_array_words + block_off + c;
        const u32 qc = (u32)q0 + c;
        const u64 jcol = a.w == 1 ? qc : qc / 3u;
        u64 *tl = tile + c * kFastS;
        if constexpr (PLAIN) {
            // asynchronous global -> shared copies (LDGSTS): all 32 rows of a lane are in flight at once and no
            // register round trip; the column pass was waiting on this staging (long_scoreboard 1.35 per issue)
            const u32 tl_s = (u32)__cvta_generic_to_shared(tl);
#pragma unroll
            for (u32 it = 0; it < 32; it++) {
                const u32 r = warp * kStageRowsPerWarp + it * kStageRowsPerIt + rsub;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tl_s + r * 8u),
                             "l"(src + r * row_words)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        }
#pragma unroll 8
        for (u32 it = 0; it < (PLAIN ? 0u : 32u); it++) {
            const u32 r = warp * kStageRowsPerWarp + it * kStageRowsPerIt + rsub;
            if constexpr (PLAIN) {
                tl[r] = src[r * row_words];
            } else {
                const u64 j = ((u64)o * 1024 + r) * inner_elems + jcol;
                u64 x = 0;
                if (j < a.n_in_elems) {
                    x = src[r * row_words];
                    if (a.pre.lo) x = gl_mul(x, scale_factor(a.pre, j));
                }
                tl[r] = x;
            }
        }
    }
    __syncthreads();

    // ---- the column of this warp ----
    u64 *slice = tile + warp * kFastS;
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) v[aa] = slice[32 * aa + lane];
    const u32 qw = (u32)q0 + warp;  // < inner_words <= 2^21; w is 1 or 3: no run-time division
    const u64 jrest = a.w == 1 ? qw : qw / 3u;
    // inter-pass twiddle omega_B^(i * j_rest), i = lane + 32 k2: straight from the full table when there is one
    dft1024_warp<INV, PLAIN ? TF21_COL_MASKMUL : true, TF21_SHL_COL>(v, slice, a.t1 + lane,
                                          (PLAIN || a.tw_full) ? a.tw_full + jrest * 1024 + lane : nullptr, lane);
    if (!PLAIN && !a.tw_full) {
        const u64 bmask = (1ull << a.log_b) - 1;
#pragma unroll 4
        for (int k2 = 0; k2 < 32; k2++) {
            const u64 e = ((u64)(lane + 32 * k2) * jrest) & bmask;
            u64 t = scale_factor(a.tw, e);
            if (a.tw_scalar) t = gl_mul(t, a.tw_scalar);
            slice[lane + 32 * k2] = gl_mul(slice[lane + 32 * k2], t);
        }
    }
    __syncthreads();

    // ---- stage out (lazy values: the next pass accepts any representative) ----
#if TF21_COL_STORE128
    if (kFastCols == 4) {  // two adjacent word-columns per lane: one 16-byte store per row
        const u32 cp = lane & 1, rs = lane >> 1;
        const u64 *t0 = tile + (2 * cp) * kFastS, *t1c = t0 + kFastS;
        ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(a.dst + (u64)b * a.dst_array_words + block_off + 2 * cp);
        const u32 row2 = row_words >> 1;  // inner_words = 1024 w: even
#pragma unroll 8
        for (u32 it = 0; it < 16; it++) {
            const u32 r = warp * kStageRowsPerWarp + it * 16 + rs;
            d2[r * row2] = make_ulonglong2(t0[r], t1c[r]);
        }
        return;
    }
#endif
    {
        u64 *dst = a.dst + (u64)b * a.dst_array_words + block_off + c;
        const u64 *tl = tile + c * kFastS;
#pragma unroll 8
        for (u32 it = 0; it < 32; it++) {
            const u32 r = warp * kStageRowsPerWarp + it * kStageRowsPerIt + rsub;
            dst[r * row_words] = tl[r];
        }
    }
}

// ---- column pass with TMA staging (the batched 2^20 transform and every PLAIN column pass) ----------------------
struct ColTmaArgs {
    const u64 *t1;       // [32][32] omega_1024^(k1 b)
    const u64 *tw_full;  // [inner_elems][1024] inter-pass twiddles (times n^-1 for the inverse first pass)
    u32 w;               // 1 | 3
    u32 slab0;           // first slab (array x outer block) of this launch: grid.y is limited to 65535
};

constexpr u32 kTmaTileWords = 1024 * kTmaTileCols;                       // 32 KiB
#ifndef TF21_TMA_MIN_BLOCKS
#define TF21_TMA_MIN_BLOCKS 5  /* 5 CTAs of 128 threads per SM (96 registers, no spills): 2.418 -> 2.390 ms per batch; 6 (80 registers) spills: 2.66 ms */
#endif
constexpr u32 kTmaMinBlocks = TF21_TMA_MIN_BLOCKS;
constexpr size_t kColTmaSmem = kFastSmem + 1024 /* alignment slack */ + 16 /* mbarrier */;

// Same arithmetic as ntt1024_col_kernel<INV, true>; the tile travels by TMA.  Shared memory: one region that is
// first the landing zone of the four box loads ([1024 rows][4 words], SWIZZLE_32B), then -- after every warp has
// pulled its column into registers -- the four private transpose slices, and finally the outgoing tile.
// TW = false: the inter-pass twiddles are left to the load of the row pass (ntt1024_row_tma_kernel<.., true>: the
// table omega_2^20^(i j) is symmetric, so the row pass reads it coalesced next to its data and both latencies overlap).
template <bool INV, bool TW>
__global__ void __launch_bounds__(kFastThreads, kTmaMinBlocks)
    ntt1024_col_tma_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map,
                           const ColTmaArgs a) {
    static_assert(kFastCols == kTmaTileCols, "one warp per word-column of the tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the swizzle pattern is a function of the absolute shared-memory address: align the tile to 1 KiB
    // (offset arithmetic on the shared array, so that the accesses stay LDS / STS rather than generic)
    u64 *tile = reinterpret_cast<u64 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    u64 *bar = tile + kFastCols * kFastS;  // 8-byte aligned, behind the private slices
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (tiles fastest: CTAs resident together read adjacent 32-byte segments of the same rows.  Slabs fastest -- so that
    // resident CTAs share their inter-pass twiddle rows in L1 -- measured 1.57 ms against 1.31 ms: DRAM locality wins)
    const u32 ct = blockIdx.x, slab = a.slab0 + blockIdx.y;
    if (tid == 0) {
        tma_prefetch_map(&src_map);
        tma_prefetch_map(&dst_map);
        mbar_init(bar, 1);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, kTmaTileWords * 8);
#pragma unroll
        for (u32 q = 0; q < 1024 / kTmaBoxRows; q++)
            tma_load_3d(tile + q * kTmaBoxRows * kTmaTileCols, &src_map, ct * kTmaTileCols, q * kTmaBoxRows, slab, bar);
        // (prefetching the tile of the CTA that will take this one's place into L2 with cp.async.bulk.prefetch.tensor,
        // 148 ... 2960 CTAs ahead: 1.277 ... 1.313 ms against 1.278 ms -- the wait on the box loads is covered by the
        // other resident CTAs)
    }
    // word (row 32 a + lane, column warp) = tile[off0 + 128 a]: the swizzle bit depends on the lane only
    const u32 off0 = tma_tile_word(lane, warp);
    const u32 qw = ct * kTmaTileCols + warp;
    const u64 jrest = a.w == 1 ? qw : qw / 3u;
#if TF21_TW_PREFETCH
    // the warp's inter-pass twiddle row (8 KiB, an L2 hit ~0.3 us away) is needed right after the second 32-point
    // step: pull it towards the SM now, two 128-byte lines per lane
    if (TW) {
        const char *twp = reinterpret_cast<const char *>(a.tw_full + jrest * 1024) + lane * 128;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(twp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(twp + 4096));
    }
#endif
    mbar_wait(bar, 0);
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) v[aa] = tile[off0 + 128 * aa];
    __syncthreads();  // the landing zone is dead: it becomes the private slices
    dft1024_warp<INV, TF21_COL_MASKMUL, TF21_SHL_COL + 16 * TF21_FIXW_COL, true>(v, tile + warp * kFastS, a.t1 + lane,
                                                            TW ? a.tw_full + jrest * 1024 + lane : nullptr, lane,
                                                            tile + off0, 128u);
    fence_proxy_async_smem();  // my generic-proxy writes of the outgoing tile -> visible to the TMA engine
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (u32 q = 0; q < 1024 / kTmaBoxRows; q++)
            tma_store_3d(&dst_map, ct * kTmaTileCols, q * kTmaBoxRows, slab, tile + q * kTmaBoxRows * kTmaTileCols);
        tma_store_commit();
        tma_store_wait_read();  // shared memory must stay valid until the engine has read it
    }
}

// diagnostic (tests): land one tile by TMA and copy the raw shared-memory image out, to pin the swizzle formula
__global__ void __launch_bounds__(128) tma_tile_probe_kernel(const __grid_constant__ CUtensorMap src_map, u64 *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *tile = reinterpret_cast<u64 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    u64 *bar = tile + kTmaTileWords;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, kTmaTileWords * 8);
        for (u32 q = 0; q < 1024 / kTmaBoxRows; q++)
            tma_load_3d(tile + q * kTmaBoxRows * kTmaTileCols, &src_map, blockIdx.x * kTmaTileCols, q * kTmaBoxRows, 0, bar);
    }
    mbar_wait(bar, 0);
    for (u32 i = threadIdx.x; i < kTmaTileWords; i += blockDim.x) out[(u64)blockIdx.x * kTmaTileWords + i] = tile[i];
}

struct FastRowArgs {
    const u64 *src;
    u64 *dst;
    u64 array_words;      // n * w
    u32 w;
    u32 n1, n2, n3;       // sizes of the leading digits i_1, i_2, i_3 (1 when absent); rows = n1 n2 n3
    u32 log_n1, log_n2;   // their logarithms (powers of two)
    u64 n_cols_total;     // batch * rows * w word-columns over the whole batch
    u32 n_tiles;          // rows * w / 8 when that is exact (a CTA never straddles arrays), else 0
    const u64 *t1;
    u64 post_scalar;      // 0 => none
    ScaleTab post;
    u32 store128;         // host decision: dst and every array of the batch 16-byte aligned (C ABI promises 8 only)
};

// Last pass of a multi-pass transform.  The passes before it are position preserving, so the source
// is [i_1][i_2][i_3][j_k][w] with contiguous 1024-element rows; the output index is
// o' + rows * i_k with o' = i_1 + n1 (i_2 + n2 i_3) (digit reversal of the row index).  A CTA takes 8
// consecutive word-columns tc = o' * w + c of the output; a warp loads its row directly (stride w),
// the transposed store goes through shared memory; canonical on store.
// POST: a scalar and / or a coset scale table is applied on store (unscale of short inverse transforms,
// fast_coset_interpolate); without it the stage-out loop has no uniform branches.
template <bool INV, u32 W, bool POST>
__global__ void __launch_bounds__(kFastThreads, kFastMinBlocks) ntt1024_row_kernel(const FastRowArgs a) {
    extern __shared__ __align__(16) u64 smem[];
    u64 *tile = smem;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 rows = a.n1 * a.n2 * a.n3;
    const u32 rw = rows * W;  // word-columns per array
    // word-column g of the whole batch: array b = g / rw, column tc = g % rw.  A CTA takes kFastCols consecutive g.
    // rows * W a multiple of kFastCols (every n >= 2^12): grid = (tiles per array, arrays), a CTA stays inside one
    // array and needs no division; otherwise (n = 2^11, XFE 2^11) the columns of a CTA may span several arrays.
    const bool tiled = a.n_tiles != 0;
    const u64 g0 = tiled ? 0 : (u64)blockIdx.x * kFastCols;
    const u64 b_cta = tiled ? blockIdx.y : 0;
    const u32 tc_cta = tiled ? blockIdx.x * kFastCols : 0;
    {
        const u64 g = g0 + warp;
        if (tiled || g < a.n_cols_total) {
            const u64 b = tiled ? b_cta : g / rw;
            const u32 tc = tiled ? tc_cta + warp : (u32)(g - b * rw);
            const u32 op = tc / W, c = tc - op * W;
            // n1, n2, n3 are powers of two (the sizes of the leading passes)
            const u32 i1 = op & (a.n1 - 1), r23 = op >> a.log_n1;
            const u32 i2 = r23 & (a.n2 - 1), i3 = r23 >> a.log_n2;
            const u32 rho = (i1 * a.n2 + i2) * a.n3 + i3;  // < rows
            const u64 *row = a.src + b * a.array_words + (u64)rho * (1024 * W) + c;
            u64 v[32];
#pragma unroll
            for (int aa = 0; aa < 32; aa++) v[aa] = row[(32 * aa + lane) * W];
            dft1024_warp<INV, false, TF21_SHL_ROW>(v, tile + warp * kFastS, a.t1 + lane, nullptr, lane);
        }
    }
    __syncthreads();

    // stage out: word (i_k = r, tc) -> dst[r * rows * w + tc]
    {
        const u32 c8 = lane % kFastCols, rsub = lane / kFastCols;
        const u64 g = g0 + c8;
        if (!tiled && g >= a.n_cols_total) return;
        const u64 b = tiled ? b_cta : g / rw;
        const u32 tc = tiled ? tc_cta + c8 : (u32)(g - b * rw);
        u64 *dst = a.dst + b * a.array_words + tc;
        const u64 *tl = tile + c8 * kFastS;
        const u32 op = tc / W;  // element index = op + rows * r
#if TF21_ROW_STORE128
        if (a.store128 && tiled && !POST && kFastCols == 4 && (rw & 1) == 0 && a.array_words < (1ull << 29)) {
            // two adjacent word-columns per lane: one 16-byte store per row pair instead of two 8-byte stores
            const u32 cp = lane & 1, rs = lane >> 1;  // 16 rows per warp instruction
            const u64 *t0 = tile + (2 * cp) * kFastS, *t1c = t0 + kFastS;
            ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(a.dst + b_cta * a.array_words + tc_cta + 2 * cp);
            const u32 ostride2 = rw >> 1;
#pragma unroll 8
            for (u32 it = 0; it < 16; it++) {
                const u32 r = warp * kStageRowsPerWarp + it * 16 + rs;
                d2[r * ostride2] = make_ulonglong2(gl_canonw(t0[r]), gl_canonw(t1c[r]));
            }
            return;
        }
#endif
        if (a.array_words < (1ull << 29)) {
            // byte offsets of an array below 4 GiB: 32-bit row offsets (one narrow IMAD per row)
            const u32 ostride = rw;
#pragma unroll 8
            for (u32 it = 0; it < 32; it++) {
                const u32 r = warp * kStageRowsPerWarp + it * kStageRowsPerIt + rsub;
                u64 x = tl[r];
                if constexpr (POST) {
                    if (a.post_scalar) x = gl_mul(x, a.post_scalar);
                    if (a.post.lo) x = gl_mul(x, scale_factor_l(a.post, op + (u64)rows * r));
                }
                dst[r * ostride] = gl_canonw(x);
            }
        } else {
            const u64 ostride = rw;
#pragma unroll 8
            for (u32 it = 0; it < 32; it++) {
                const u32 r = warp * kStageRowsPerWarp + it * kStageRowsPerIt + rsub;
                u64 x = tl[r];
                if constexpr (POST) {
                    if (a.post_scalar) x = gl_mul(x, a.post_scalar);
                    if (a.post.lo) x = gl_mul(x, scale_factor_l(a.post, op + (u64)rows * r));
                }
                dst[(u64)r * ostride] = gl_canonw(x);
            }
        }
    }
}

// ---- last (transposing) pass with a TMA store ----------------------------------------------------------------
struct RowTmaArgs {
    const u64 *src;
    u64 array_words;
    u32 n1, n2, n3, log_n1, log_n2;
    const u64 *t1;
    u32 b0;  // first array of this launch (grid.y <= 65535)
    const u64 *tw_in;  // TWIN: [1024][1024] inter-pass twiddles of the preceding column pass, applied on load
};

// ntt1024_row_kernel<INV, W, false> for tiled shapes with an aligned destination: a warp loads its contiguous row,
// the second 32-point step writes canonical words into the [1024 rows][4 word-columns] tile in TMA layout and one
// thread sends it to dst viewed as [batch][1024][rows * W] with four box stores.
template <bool INV, u32 W, bool TWIN>
__global__ void __launch_bounds__(kFastThreads, kTmaMinBlocks)
    ntt1024_row_tma_kernel(const __grid_constant__ CUtensorMap dst_map, const RowTmaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *tile = reinterpret_cast<u64 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 tc_cta = blockIdx.x * kFastCols, b = a.b0 + blockIdx.y;
    if (tid == 0) tma_prefetch_map(&dst_map);
    const u32 tc = tc_cta + warp;
    const u32 op = tc / W, c = tc - op * W;
    const u32 i1 = op & (a.n1 - 1), r23 = op >> a.log_n1;
    const u32 i2 = r23 & (a.n2 - 1), i3 = r23 >> a.log_n2;
    const u32 rho = (i1 * a.n2 + i2) * a.n3 + i3;
    const u64 *row = a.src + (u64)b * a.array_words + (u64)rho * (1024 * W) + c;
    u64 v[32];
    if constexpr (TWIN) {
        // element j of row rho of a 2^20-point block carries omega^(i j) with i = rho mod 1024 (the column-pass digit
        // is the last leading digit); T[i][j] = T[j][i]
        const u64 *twr = a.tw_in + (u64)(rho & 1023u) * 1024 + lane;
#pragma unroll
        for (int aa = 0; aa < 32; aa++) v[aa] = gl_mul(row[(32 * aa + lane) * W], __ldg(twr + 32 * aa));
    } else {
#pragma unroll
        for (int aa = 0; aa < 32; aa++) v[aa] = row[(32 * aa + lane) * W];
    }
    const u32 off0 = tma_tile_word(lane, warp);
    dft1024_warp<INV, false, TF21_SHL_ROW + 16 * TF21_FIXW_ROW, true, true>(v, tile + warp * kFastS, a.t1 + lane, nullptr, lane, tile + off0, 128u);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (u32 q = 0; q < 1024 / kTmaBoxRows; q++)
            tma_store_3d(&dst_map, tc_cta, q * kTmaBoxRows, b, tile + q * kTmaBoxRows * kTmaTileCols);
        tma_store_commit();
        tma_store_wait_read();
    }
}

struct SmallColArgs {
    const u64 *src;
    u64 *dst;
    u64 src_array_words, dst_array_words;
    u32 w;
    u64 inner_words;     // multiple of 128
    u32 n_outer;
    u64 n_in_elems;
    const u64 *tw_full;  // [N_p][inner_elems] scalar * omega_B^(i j) (B <= 2^20) or nullptr
    ScaleTab tw;         // split tables for omega_B^e, used when tw_full == nullptr
    u64 tw_scalar;       // extra factor on every output of the split path (n^-1), 0 => none
    ScaleTab pre;
};

// Column pass of 2^A <= 64 points: one THREAD owns one word-column (adjacent threads own adjacent
// columns, so every access is fully coalesced and nothing goes through shared memory); the whole
// transform uses shift twiddles; then the inter-pass twiddle omega_B^(i * j_rest).
template <bool INV, int A>
__global__ void TF21_SMALLCOL_BOUNDS ntt_small_col_kernel(const SmallColArgs a) {
    constexpr int NP = 1 << A;
    const u64 gid = (u64)blockIdx.x * 128 + threadIdx.x;
    const u64 q = gid % a.inner_words;
    const u64 rest = gid / a.inner_words;
    const u32 o = (u32)(rest % a.n_outer);
    const u64 b = rest / a.n_outer;
    const u64 inner_elems = a.inner_words / a.w;
    const u64 jcol = q / a.w;
    const u64 off = (u64)o * NP * a.inner_words + q;
    const u64 *src = a.src + b * a.src_array_words + off;
    u64 v[NP];
#pragma unroll
    for (int r = 0; r < NP; r++) {
        const u64 j = ((u64)o * NP + r) * inner_elems + jcol;
        u64 x = 0;
        if (j < a.n_in_elems) {
            x = src[(u64)r * a.inner_words];
            if (a.pre.lo) x = gl_mul(x, scale_factor_l(a.pre, j));
        }
        v[r] = x;
    }
    dft_pow2<INV, A>(v);
    u64 *dst = a.dst + b * a.dst_array_words + off;
    if (a.tw_full) {
        const u64 *tcol = a.tw_full + jcol;  // consecutive threads read consecutive entries
#pragma unroll
        for (int i = 0; i < NP; i++)
            dst[(u64)i * a.inner_words] = gl_mul(v[brev_bits(i, A)], __ldg(tcol + (u64)i * inner_elems));
    } else {
        // B > 2^20: no full table.  g = omega_B^jcol from the split tables (coalesced: the index is the
        // column), then scalar * g^i by four interleaved power chains -- no gathers.
        const u64 g = scale_factor_l(a.tw, jcol);
        u64 pw[4];
        pw[0] = a.tw_scalar ? a.tw_scalar : 1ull;
        pw[1] = gl_mul(pw[0], g);
        pw[2] = gl_mul(pw[1], g);
        pw[3] = gl_mul(pw[2], g);
        const u64 g2 = gl_mul(g, g);
        const u64 g4 = gl_mul(g2, g2);
#pragma unroll
        for (int i = 0; i < NP; i++) {
            if (i >= 4) pw[i & 3] = gl_mul(pw[i & 3], g4);
            dst[(u64)i * a.inner_words] = gl_mul(v[brev_bits(i, A)], pw[i & 3]);
        }
    }
}

// omega_64^k, k < 64 (canonical; = +-2^(39 k mod 96)), filled by get_tables: the twiddles of the pruned pass below,
// read with a warp-uniform index (constant-bank broadcast)
__constant__ u64 c_w64[64];

// Pruned column pass for zero-extended inputs (the `resize(order, ZERO)` of fast_coset_evaluate,
// polynomial.rs:1396-1397): only the first NZ = 2^LNZ rows of a 2^A-point column can be non-zero, so
// with M = 2^A / NZ and k = M c + d:  X[M c + d] = sum_{a < NZ} (x_a w^(a d)) w_NZ^(a c)  -- for every d one
// NZ-point DFT of twiddled inputs.  Reads NZ rows, writes 2^A rows, keeps 2 NZ values live.
// The loop over d stays ROLLED: unrolled (shift twiddles as immediates) the 64-point form is 74 KB of
// straight-line code and ran at 6.9 `no_instruction` stalls per issue (profiles/r02d_ncu_lde26_summary.txt);
// the twiddles w^(a d) are now general products with a warp-uniform table entry.
template <int A, int LNZ>
__global__ void TF21_PRUNED_BOUNDS ntt_small_col_pruned_kernel(const SmallColArgs a) {
    constexpr int NP = 1 << A, NZ = 1 << LNZ, M = NP / NZ;
    const u64 gid = (u64)blockIdx.x * 128 + threadIdx.x;
    const u64 q = gid % a.inner_words;
    const u64 rest = gid / a.inner_words;
    const u32 o = (u32)(rest % a.n_outer);
    const u64 b = rest / a.n_outer;
    const u64 inner_elems = a.inner_words / a.w;
    const u64 jcol = q / a.w;
    const u64 off = (u64)o * NP * a.inner_words + q;
    const u64 *src = a.src + b * a.src_array_words + off;
    u64 x[NZ];
#pragma unroll
    for (int r = 0; r < NZ; r++) {
        const u64 j = ((u64)o * NP + r) * inner_elems + jcol;
        u64 v = 0;
        if (j < a.n_in_elems) {
            v = src[(u64)r * a.inner_words];
            if (a.pre.lo) v = gl_mul(v, scale_factor_l(a.pre, j));
        }
        x[r] = v;
    }
    u64 *dst = a.dst + b * a.dst_array_words + off;
    if (!a.tw_full) {
        // B > 2^20: no table.  With g = omega_B^jcol the inter-pass twiddle of output i = M c + d is g^d (g^M)^c, and g^d
        // can be carried by the inputs:  out[M c + d] = (g^M)^c  sum_a (x_a h_a^d) w_NZ^(a c),  h_a = w^a g.
        // The running values x_a h_a^d take one product per input and d, the outputs one product for c > 0:
        // NZ + NZ - 1 products per NZ outputs (the first form -- w^(a d) from a table, a g^i chain and the product
        // with it -- needed 3 NZ - 1).
        const u64 g = scale_factor_l(a.tw, jcol);
        u64 gM = g;
#pragma unroll
        for (int k = 0; k < A - LNZ; k++) gM = gl_mul(gM, gM);
        u64 h[NZ], P[NZ];
        h[0] = g;
        P[0] = 1;
#pragma unroll
        for (int r = 1; r < NZ; r++) {
            h[r] = gl_mul(g, c_w64[r * (64 / NP)]);
            P[r] = r == 1 ? gM : gl_mul(P[r - 1], gM);
        }
#pragma unroll 1
        for (int d = 0; d < M; d++) {
            u64 z[NZ];
#pragma unroll
            for (int r = 0; r < NZ; r++) z[r] = x[r];
            dft_pow2<false, LNZ>(z);
#pragma unroll
            for (int c = 0; c < NZ; c++) {
                const u64 v = z[brev_bits(c, LNZ)];
                dst[(u64)(M * c + d) * a.inner_words] = c ? gl_mul(v, P[c]) : v;  // lazy words: the next pass takes any
            }
#pragma unroll
            for (int r = 0; r < NZ; r++) x[r] = gl_mul(x[r], h[r]);
        }
        return;
    }
    // B <= 2^20: the twiddles omega_B^(i jcol) come from the full table (consecutive threads read consecutive entries)
    const u64 *tcol = a.tw_full + jcol;
#pragma unroll 1
    for (int d = 0; d < M; d++) {
        u64 z[NZ];
        z[0] = x[0];
#pragma unroll
        for (int r = 1; r < NZ; r++) z[r] = gl_mul(x[r], c_w64[((r * d) & (NP - 1)) * (64 / NP)]);
        dft_pow2<false, LNZ>(z);
#pragma unroll
        for (int c = 0; c < NZ; c++) {
            const int i = M * c + d;
            dst[(u64)i * a.inner_words] = gl_mul(z[brev_bits(c, LNZ)], __ldg(tcol + (u64)i * inner_elems));
        }
    }
}

struct FastSingleArgs {
    const u64 *src;
    u64 *dst;
    u64 src_array_words, dst_array_words;
    u64 batch;
    u64 n_in_elems;
    const u64 *t1;
    u64 post_scalar;
    ScaleTab pre, post;
};

// n = 1024: one warp per word-column (array, c) of the batch, no global staging at all.  For w = 3 the three
// warps of an array interleave their 8-byte accesses (stride 24 B): every sector is still used completely.
template <bool INV, u32 W>
__global__ void __launch_bounds__(kFastThreads, kFastMinBlocks) ntt1024_single_kernel(const FastSingleArgs a) {
    extern __shared__ __align__(16) u64 smem[];
    u64 *tile = smem;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 g = (u64)blockIdx.x * kFastCols + warp;
    if (g >= a.batch * W) return;
    const u64 arr = g / W;
    const u32 c = (u32)(g - arr * W);
    const u64 *row = a.src + arr * a.src_array_words + c;
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) {
        const u64 j = 32 * aa + lane;
        u64 x = 0;
        if (j < a.n_in_elems) {
            x = row[j * W];
            if (a.pre.lo) x = gl_mul(x, scale_factor(a.pre, j));
        }
        v[aa] = x;
    }
    u64 *slice = tile + warp * kFastS;
    dft1024_warp<INV, false, TF21_SHL_SINGLE>(v, slice, a.t1 + lane, nullptr, lane);
    u64 *out = a.dst + arr * a.dst_array_words + c;
#pragma unroll 8
    for (int k2 = 0; k2 < 32; k2++) {
        u64 x = slice[lane + 32 * k2];
        if (a.post_scalar) x = gl_mul(x, a.post_scalar);
        if (a.post.lo) x = gl_mul(x, scale_factor(a.post, (u64)(lane + 32 * k2)));
        out[(u64)(lane + 32 * k2) * W] = gl_canonw(x);
    }
}


// ---- leading pass of 2^K points, 1 <= K <= 9, in registers (n = 2^11 .. 2^19 and the top digit of n >= 2^21) ----
// Position preserving like every leading pass: [outer][2^K rows][inner], DFT along the rows, then the twiddle
// omega_B^(kappa * j_rest).  A warp takes a [2^K rows][G = 2^(10-K) word-columns] tile = 1024 elements, flat index
// e = row * G + col = 32 a + lane:
//   K < 5 : the rows are the high K bits of the register index a -- strided register groups, no shared memory;
//   K >= 5: step A = 32-point DFT over a, twiddle omega_2^K^(kappa1 * b_hi), 32 x 32 transpose, step B =
//           2^(K-5)-point DFTs over the high bits of b (register groups of stride G); X[kappa1 + 32 kappa2] of
//           column b_lo sits in lane kappa1; a second pass through the slice restores e-order for coalesced stores.
// Replaces two thread-per-column passes (<= 32 points each) for 2^16 .. 2^19 and 2^26 .. 2^29.
struct ColNArgs {
    const u64 *src;
    u64 *dst;
    u64 array_words;    // n * w: distance between arrays (source and destination: plain passes only)
    u32 inner_words;    // row stride in words (inner_elems * w)
    u32 w;
    u32 n_outer;
    const u64 *tw1;     // K > 5: [32][2^(K-5)] omega_2^K^(+-kappa1 b_hi); else nullptr
    const u64 *tw_full; // [2^K][inner_elems] scalar * omega_B^(+-kappa j) or nullptr (then the split tables)
    ScaleTab tw;
    u32 log_b;
    u32 scaled;         // the full table carries a scalar other than one (its row 0 is not all ones)
};

template <bool INV, int K>
__global__ void __launch_bounds__(kFastThreads, (K <= 2 ? 5 : kFastMinBlocks)) ntt_col_n_kernel(const ColNArgs a) {
    static_assert(K >= 1 && K <= 9, "leading passes of 2 .. 512 points");
    constexpr u32 G = 1u << (10 - K);  // word-columns per warp
    extern __shared__ __align__(16) u64 smem[];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 col0 = (blockIdx.x * kFastCols + warp) * G;
    if (col0 >= a.inner_words) return;
    const u32 o = blockIdx.y, b = blockIdx.z;
    const u64 base = (u64)b * a.array_words + (u64)o * ((u64)a.inner_words << K) + col0;
    const u64 *src = a.src + base;
    u64 *dst = a.dst + base;
    const u32 one = c_gl_one;
    u64 v[32];
    // element e = 32 aa + lane: row e >> (10 - K), column e & (G - 1)
#pragma unroll
    for (int aa = 0; aa < 32; aa++) {
        const u32 e = 32 * aa + lane;
        v[aa] = __ldcs(src + (u64)(e >> (10 - K)) * a.inner_words + (e & (G - 1)));
    }
    const u64 bmask = (1ull << a.log_b) - 1;
    const u32 inner_elems = a.w == 1 ? a.inner_words : a.inner_words / 3u;
    if constexpr (K <= 5) {
        dft_groups_strided<INV, K>(v, one);
        // register brev_K(kappa) * 2^(5-K) + a_lo holds X[kappa] of column a_lo * 32 + lane: coalesced as it is
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const u32 kappa = brev_bits((u32)r >> (5 - K), K), a_lo = (u32)r & ((1u << (5 - K)) - 1);
            const u32 col = a_lo * 32 + lane;
            const u32 jrest = a.w == 1 ? col0 + col : (col0 + col) / 3u;
            u64 x = v[r];
            if (a.tw_full) {
                if (kappa != 0 || a.scaled) x = gl_mul(x, __ldg(a.tw_full + (u64)kappa * inner_elems + jrest));
            } else if (kappa != 0) {
                x = gl_mul(x, scale_factor_l(a.tw, ((u64)kappa * jrest) & bmask));
            }
            __stcs(dst + (u64)kappa * a.inner_words + col, x);
        }
    } else {
        constexpr int BP = K - 5;  // bits of b_hi
        u64 *slice = smem + warp * kFastS;
        dft_groups<INV, 5>(v, one);
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) {
            u64 x = v[brev5((u32)k1)];
            if (k1 != 0) x = gl_mul(x, __ldg(a.tw1 + (k1 << BP) + (lane >> (10 - K))));
            slice[k1 * kTransposeStride + lane] = x;
        }
        __syncwarp();
#pragma unroll
        for (int bb = 0; bb < 32; bb++) v[bb] = slice[lane * kTransposeStride + bb];
        dft_groups_strided<INV, BP>(v, one);
        __syncwarp();
        // lane = kappa1; register brev_BP(kappa2) * G + b_lo holds X[kappa1 + 32 kappa2] of column b_lo -> e-order.
        // One pad word per 16 (slot(e) = e + e / 16): the writes of a half warp (stride G words) and the reads
        // (consecutive e) both fall on 16 different bank pairs.
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const u32 kappa2 = brev_bits((u32)r >> (10 - K), BP), b_lo = (u32)r & (G - 1);
            const u32 e = (lane + 32 * kappa2) * G + b_lo;
            slice[e + (e >> 4)] = v[r];
        }
        __syncwarp();
#pragma unroll 8
        for (int aa = 0; aa < 32; aa++) {
            const u32 e = 32 * aa + lane;
            const u32 kappa = e >> (10 - K), col = e & (G - 1);
            const u32 jrest = a.w == 1 ? col0 + col : (col0 + col) / 3u;
            u64 x = slice[e + (e >> 4)];
            if (a.tw_full) x = gl_mul(x, __ldg(a.tw_full + (u64)kappa * inner_elems + jrest));
            else x = gl_mul(x, scale_factor_l(a.tw, ((u64)kappa * jrest) & bmask));
            __stcs(dst + (u64)kappa * a.inner_words + col, x);
        }
    }
}

// ---- leading pass of 2^K points, 5 <= K <= 9, one CTA per [2^K rows][C word-columns] tile ----------------------------
// K = A + B, A = ceil(K / 2).  256 threads; thread (c, g): c = tid % C (fastest: a warp instruction touches 32 / 16
// consecutive words of one or two rows), g = tid / C < 2^B, C = 256 / 2^B.
//   step 1: rows a 2^B + g (a < 2^A) of word-column c in registers: 2^A-point DFT over a (shift twiddles) -> k1,
//           times omega_{2^K}^(k1 g), into shared memory s[(k1 2^B + g) C + c];
//   step 2: thread (c, q) takes k1 = q + 2^B h (h < 2^(A-B)): 2^B-point DFT over g -> k2; row kappa = k1 + 2^A k2 gets the
//           inter-pass twiddle and goes back to its own position (position preserving, like every leading pass).
// Against the passes it replaces: one pass over the data instead of two thread-per-column passes of <= 32 points
// (2^16 .. 2^19 and 2^26 .. 2^29), 8 .. 32 elements per thread instead of 32 .. 64 (more resident warps than the
// warp-per-tile form ntt_col_n_kernel, which is latency bound), row segments of 128 .. 512 bytes.
constexpr u32 kMidThreads = 256;
// resident CTAs per SM the register allocation aims for: the pass is bound by the exposed latency of its one load phase
// per CTA, so resident warps count for more than spill-free code (profiles/r02k_ab_mid_col_v4_resident_ctas.txt, ms per
// GiB pass + row pass: 4 / 3 / 2 CTAs: 2^16 1.136, 2^17 1.241, 2^18 1.512; 5 / 4 / 3: 1.093, 1.209, 1.313; 6 / 5 / 4: 1.092,
// 1.246, 1.283).  An L2 prefetch of a later tile's rows made it slower (+5 %, ..._v5_l2_prefetch.txt).
#ifndef TF21_MID_BLOCKS_LO
#define TF21_MID_BLOCKS_LO 5
#endif
#ifndef TF21_MID_BLOCKS_78
#define TF21_MID_BLOCKS_78 4
#endif
#ifndef TF21_MID_OPT_CANON
#define TF21_MID_OPT_CANON TF21_OPT_CANON
#endif
// whole warps only (every thread of the CTA is active): the optimistic stages of dft32_warp for any 2^A
template <bool INV, int A>
__device__ __forceinline__ void mid_dft(u64 (&v)[1 << A], u32 one) {
#if TF21_MID_OPT_CANON
    dft_opt_stage<INV, A, 1, TF21_SHL_WIDE>(v, one);
#else
    dft_pow2<INV, A>(v, one);
#endif
}
template <int K>
struct MidShape {
    static constexpr int A = (K + 1) / 2, B = K - A, NA = 1 << A, NB = 1 << B, H = 1 << (A - B);
    static constexpr u32 C = kMidThreads >> B;
    static constexpr size_t smem = ((size_t)C << K) * sizeof(u64);
};

// (A persistent, double-buffered form -- the next tile arriving by 16-byte cp.async copies while the current one is
// transformed in place -- measured slower: 1.28 against 1.14 ms per GiB at 2^16, profiles/r02k_ab_mid_col_v3_pipelined.txt;
// the extra shared-memory round trip and the per-tile index arithmetic cost more than the exposed load latency of
// four independent resident CTAs.)
template <bool INV, int K>
__global__ void __launch_bounds__(kMidThreads, (K <= 6 ? TF21_MID_BLOCKS_LO : K <= 8 ? TF21_MID_BLOCKS_78 : 3)) ntt_mid_col_kernel(const ColNArgs a) {
    static_assert(K >= 3 && K <= 9, "leading passes of 8 .. 512 points");
    using S = MidShape<K>;
    constexpr int A = S::A, B = S::B, NA = S::NA, NB = S::NB, H = S::H;
    constexpr u32 C = S::C;
    extern __shared__ __align__(16) u64 smem[];
    const u32 tid = threadIdx.x, c = tid % C, g = tid / C;
    const u32 col = blockIdx.x * C + c;
    const u32 o = blockIdx.y, b = blockIdx.z;
    const u64 base = (u64)b * a.array_words + (u64)o * ((u64)a.inner_words << K) + col;
    const u64 *src = a.src + base;
    u64 *dst = a.dst + base;
    u64 v[NA];
#pragma unroll
    for (int aa = 0; aa < NA; aa++) v[aa] = __ldcs(src + (u64)(aa * NB + (int)g) * a.inner_words);
    const u32 one = (u32)__ldg(a.tw1);  // omega^0: an opaque 1 in a register (see gl_subp)
    mid_dft<INV, A>(v, one);
    {
        const u64 *tw1 = a.tw1 + g;
        u64 *sp = smem + g * C + c;
#pragma unroll
        for (int k1 = 0; k1 < NA; k1++) {
            u64 x = v[brev_bits((u32)k1, A)];
            if (k1 != 0) x = gl_mul(x, __ldg(tw1 + k1 * NB));
            sp[(u32)k1 * NB * C] = x;
        }
    }
    __syncthreads();
    const u64 bmask = (1ull << a.log_b) - 1;
    const u32 inner_elems = a.w == 1 ? a.inner_words : a.inner_words / 3u;
    const u32 jrest = a.w == 1 ? col : col / 3u;
#pragma unroll
    for (int h = 0; h < H; h++) {
        const u32 k1 = g + (u32)NB * (u32)h;
        u64 z[NB];
#pragma unroll
        for (int gg = 0; gg < NB; gg++) z[gg] = smem[(k1 * NB + (u32)gg) * C + c];
        mid_dft<INV, B>(z, one);
#pragma unroll
        for (int k2 = 0; k2 < NB; k2++) {
            const u32 kappa = k1 + (u32)NA * (u32)k2;
            u64 x = z[brev_bits((u32)k2, B)];
            if (a.tw_full) {
                x = gl_mul(x, __ldg(a.tw_full + (u64)kappa * inner_elems + jrest));
            } else if (kappa != 0) {
                x = gl_mul(x, scale_factor_l(a.tw, ((u64)kappa * jrest) & bmask));
            }
            __stcs(dst + (u64)kappa * a.inner_words, x);
        }
    }
}

// 512 points (K = 9) with 16 elements per thread: the 32 elements per thread of the generic form leave two resident
// CTAs per SM (1.64 against 1.57 ms per GiB for the two thread-per-column passes it would replace).  Here 512 = 16 x 32:
// step 1 as above (16-point transforms over a, times omega_512^(k1 g)); in step 2 the 32-point transform of a k1 is
// shared by TWO threads -- thread h = 0 takes the even g, h = 1 the odd g (16-point transforms E, O), the odd side applies
// omega_32^k (a shift twiddle), the halves change hands through the tile (one more barrier), and thread h forms
// X[k + 16 h] = E[k] +- omega_32^k O[k].  Tile [512 rows][8 words] (64-byte row segments), 32 KB.
// Generalised to K = 2 HB + 1 (HB = 4: 512 points, 16 elements per thread; HB = 3: 128 points, 8 elements per thread).
template <int HB>
struct Mid2Shape {
    static constexpr int E = 1 << HB, NG = 2 << HB, K = 2 * HB + 1;
    static constexpr u32 C = kMidThreads / NG;                       // word-columns per tile
    static constexpr size_t smem = ((size_t)C << K) * sizeof(u64);  // [2^K rows][C]
    static constexpr int EU = (39 << (6 - (HB + 1))) % 192;         // omega_NG = 2^EU
};
template <bool INV, int HB, int KK>
__device__ __forceinline__ void mid2_combine(u64 (&z)[1 << HB], const u64 *part, u32 h, u32 one, u64 (&x)[1 << HB]) {
    if constexpr (KK < (1 << HB)) {
        // omega_NG^k = 2^(EU k); exponents >= 96 use 2^96 = -1 and swap sum and difference
        constexpr int E0 = (Mid2Shape<HB>::EU * KK) % 192, E = INV ? (192 - E0) % 192 : E0;
        constexpr bool neg = E >= 96;
        const u64 mine = z[brev_bits((u32)KK, HB)], theirs = part[KK * Mid2Shape<HB>::C];
        // h = 0 holds E[k] and reads the twiddled O[k]; h = 1 holds the twiddled O[k] and reads E[k]
        const u64 e = h ? theirs : mine, o = h ? mine : theirs;
        const bool add = (h == 0) != neg;
        x[KK] = add ? gl_addl(e, o) : gl_subl(e, o, one);
        mid2_combine<INV, HB, KK + 1>(z, part, h, one, x);
    }
}
template <bool INV, int HB, int KK>
__device__ __forceinline__ void mid2_twiddle_odd(u64 (&z)[1 << HB], u32 one) {
    if constexpr (KK < (1 << HB)) {
        constexpr int E0 = (Mid2Shape<HB>::EU * KK) % 192, E = INV ? (192 - E0) % 192 : E0;
        constexpr int S = E >= 96 ? E - 96 : E;
        u64 &r = z[brev_bits((u32)KK, HB)];
        if constexpr (S == 0) r = gl_canonw(r);
        else r = gl_shlc<(S ? S : 1)>(r, one);
        mid2_twiddle_odd<INV, HB, KK + 1>(z, one);
    }
}
#ifndef TF21_MID2_BLOCKS_7
#define TF21_MID2_BLOCKS_7 5
#endif
template <bool INV, int HB>
__global__ void __launch_bounds__(kMidThreads, (HB == 4 ? 4 : HB == 3 ? TF21_MID2_BLOCKS_7 : 6)) ntt_mid2_col_kernel(const ColNArgs a) {
    using S = Mid2Shape<HB>;
    constexpr int E = S::E, NG = S::NG, K = S::K;
    constexpr u32 C = S::C;
    extern __shared__ __align__(16) u64 smem[];
    const u32 tid = threadIdx.x, c = tid % C, q = tid / C;  // q < NG
    const u32 col = blockIdx.x * C + c;
    const u32 o = blockIdx.y, b = blockIdx.z;
    const u64 base = (u64)b * a.array_words + (u64)o * ((u64)a.inner_words << K) + col;
    const u64 *src = a.src + base;
    u64 *dst = a.dst + base;
    const u32 one = (u32)__ldg(a.tw1);
    u64 v[E];
#pragma unroll
    for (int aa = 0; aa < E; aa++) v[aa] = __ldcs(src + (u64)(aa * NG + (int)q) * a.inner_words);
    mid_dft<INV, HB>(v, one);
    {
        const u64 *tw1 = a.tw1 + q;  // [E][NG]: omega_{2^K}^(k1 g)
        u64 *sp = smem + q * C + c;
#pragma unroll
        for (int k1 = 0; k1 < E; k1++) {
            u64 x = v[brev_bits((u32)k1, HB)];
            if (k1 != 0) x = gl_mul(x, __ldg(tw1 + k1 * NG));
            sp[(u32)k1 * NG * C] = x;
        }
    }
    __syncthreads();
    const u32 k1 = q & (u32)(E - 1), h = q >> HB;
    u64 z[E];
#pragma unroll
    for (int gg = 0; gg < E; gg++) z[gg] = smem[(k1 * NG + 2 * (u32)gg + h) * C + c];
    mid_dft<INV, HB>(z, one);
    // the odd side carries omega_NG^k as a canonical word (an addend of the lazy sums); the even side's E[k] is only ever
    // the first (lazy) operand
    if (h) mid2_twiddle_odd<INV, HB, 0>(z, one);
    __syncthreads();  // every thread has read its step-2 inputs: the tile can carry the halves
    {
        u64 *sp = smem + (k1 * NG + h * E) * C + c;
#pragma unroll
        for (int k = 0; k < E; k++) sp[(u32)k * C] = z[brev_bits((u32)k, HB)];
    }
    __syncthreads();
    u64 x[E];
    // the partner's value for index k sits at row k1 * NG + (1 - h) * E + k
    mid2_combine<INV, HB, 0>(z, smem + (k1 * NG + (1u - h) * E) * C + c, h, one, x);
    const u64 bmask = (1ull << a.log_b) - 1;
    const u32 inner_elems = a.w == 1 ? a.inner_words : a.inner_words / 3u;
    const u32 jrest = a.w == 1 ? col : col / 3u;
#pragma unroll
    for (int k = 0; k < E; k++) {
        const u32 kappa = k1 + (u32)E * ((u32)k + (u32)E * h);
        u64 y = x[k];
        if (a.tw_full) {
            y = gl_mul(y, __ldg(a.tw_full + (u64)kappa * inner_elems + jrest));
        } else if (kappa != 0) {
            y = gl_mul(y, scale_factor_l(a.tw, ((u64)kappa * jrest) & bmask));
        }
        __stcs(dst + (u64)kappa * a.inner_words, y);
    }
}

// ---- n = 2^K < 1024: a warp takes 1024 consecutive elements = 2^(10-K) whole columns, all in registers ------
// (reference benches at 2^7, benches/ntt.rs:19).  Flat element J of the batch = (array J / n, index J % n); a warp
// owns J0 .. J0 + 1023 of one coefficient lane (w = 3: the three warps of a chunk interleave like the 2^10 kernel).
//   K > 5: step A = 2^(K-5)-point DFTs over the low bits of the register index, twiddle omega_n^(k1 lane), 32 x 32
//          transpose, step B = 32-point DFT: X[k1 + 2^(K-5) k2] of column c sits in lane (c, k1), register k2;
//   K <= 5: transpose first, then 2^(5-K) independent 2^K-point DFTs on register groups (no general product).
struct SmallNArgs {
    const u64 *src;
    u64 *dst;
    u64 total_elems;   // batch * n
    const u64 *tw;     // K > 5: [2^(K-5)][32] scalar * omega_n^(+-k1 b); else nullptr
    u64 post_scalar;   // K <= 5: n^-1 applied on store (0 = none)
    u32 w;
    u32 scaled;        // K > 5: the table carries a scalar other than one (row 0 is not all ones)
};

template <bool INV, int K>
// 5 resident CTAs (96 registers, 8 bytes of spill) for K = 6, 7: 2^6 0.707 -> 0.669 ms, 2^7 0.699 -> 0.610 ms per GiB; the other
// sizes and the 2^10 single-pass kernel lose 1 .. 4 % with it (profiles/r02k_ab_min_blocks5_register_kernels.txt)
__global__ void __launch_bounds__(kFastThreads, ((K == 6 || K == 7) ? 5 : kFastMinBlocks)) ntt_small_n_kernel(const SmallNArgs a) {
    static_assert(K >= 1 && K <= 9, "sizes below 2^10");
    extern __shared__ __align__(16) u64 smem[];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 g = (u64)blockIdx.x * kFastCols + warp;  // (chunk, coefficient lane)
    const u64 chunk = g / a.w;
    const u32 c = (u32)(g - chunk * a.w);
    const u64 j0 = chunk * 1024;
    if (j0 >= a.total_elems) return;
    const u64 left = a.total_elems - j0;  // a multiple of n; < 1024 only in the last chunk
    const u64 *src = a.src + j0 * a.w + c;
    u64 *slice = smem + warp * kFastS;
    const u32 one = c_gl_one;
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) {
        const u32 j = 32 * aa + lane;
        v[aa] = j < left ? src[(u64)j * a.w] : 0ull;
    }
    u32 base;  // position of register 0 of this lane in the chunk after the last step; register k2 adds k2 * stride
    constexpr u32 stride = K > 5 ? (1u << (K - 5)) : 1u;
    if constexpr (K > 5) {
        constexpr int AP = K - 5;
        dft_groups<INV, AP>(v, one);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const int grp = r >> AP, k1 = r & ((1 << AP) - 1);
            const int reg = (grp << AP) + (int)brev_bits((u32)k1, AP);
            u64 x = v[reg];
            if (k1 != 0 || a.scaled) x = gl_mul(x, __ldg(a.tw + k1 * 32 + lane));  // unscaled row 0 is all ones
            slice[r * kTransposeStride + lane] = x;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 32; b++) v[b] = slice[lane * kTransposeStride + b];
        dft_groups<INV, 5>(v, one);
        base = ((lane >> AP) << K) + (lane & ((1u << AP) - 1));
    } else {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r++) slice[r * kTransposeStride + lane] = v[r];
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 32; b++) v[b] = slice[lane * kTransposeStride + b];
        dft_groups<INV, K>(v, one);
        base = 32 * lane;
    }
    __syncwarp();
    // natural order into the slice, then coalesced stores
#pragma unroll
    for (int r = 0; r < 32; r++) {
        // K > 5: register brev5(k2) holds X[.. + stride * k2]; K <= 5: register (grp << K) + brev_K(k) holds X[grp][k]
        const int reg = K > 5 ? (int)brev5((u32)r) : ((r >> K) << K) + (int)brev_bits((u32)(r & ((1 << K) - 1)), K);
        slice[base + (u32)r * stride] = v[reg];
    }
    __syncwarp();
    u64 *dst = a.dst + j0 * a.w + c;
#pragma unroll 8
    for (int aa = 0; aa < 32; aa++) {
        const u32 j = 32 * aa + lane;
        if (j < left) {
            u64 x = slice[j];
            if (K <= 5 && a.post_scalar) x = gl_mul(x, a.post_scalar);
            dst[(u64)j * a.w] = gl_canonw(x);
        }
    }
}

// ---- tables for the fast path ---------------------------------------------------------------------
struct FastTables {
    std::map<int, u64 *> t1;                                             // inverse -> [32][32]
    std::map<std::tuple<unsigned, unsigned, int, u64>, u64 *> tw_full;   // (log_b, log_np, inverse, scalar)
    std::map<std::tuple<unsigned, unsigned, int, u64>, u64 *> tw_small;  // same key, [N_p][B / N_p] layout
};
static std::map<int, FastTables> g_fast_tables;  // by device, guarded by g_mutex
static std::map<std::tuple<int, unsigned, int, u64>, u64 *> g_small_n_tw;  // (device, K, inverse, scalar), guarded by g_mutex

static std::map<std::tuple<int, unsigned, int>, u64 *> g_col_n_tw1;  // (device, K, inverse), guarded by g_mutex

// [32][2^(K-5)] omega_{2^K}^(+-kappa1 b_hi) for the 2^K-point leading pass, 5 < K < 10
inline int get_col_n_tw1(DeviceTables &t, int dev, unsigned k, int inverse, const u64 **out) {
    auto key = std::make_tuple(dev, k, inverse);
    auto it = g_col_n_tw1.find(key);
    if (it != g_col_n_tw1.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(k);
    if (inverse) w = hgl_inv(w);
    const u32 cols = 1u << (k - 5);
    std::vector<u64> h(32 * cols);
    for (u32 k1 = 0; k1 < 32; k1++) {
        u64 step = hgl_pow(w, k1), acc = 1;
        for (u32 bh = 0; bh < cols; bh++) {
            h[k1 * cols + bh] = acc;
            acc = hgl_mul(acc, step);
        }
    }
    u64 *d;
    TF21_TRY(upload(t, h, &d));
    g_col_n_tw1[key] = d;
    *out = d;
    return 0;
}

static std::map<std::tuple<int, unsigned, int>, u64 *> g_mid_tw1;  // (device, K, inverse), guarded by g_mutex

// [2^A][2^B] omega_{2^K}^(+-k1 g) for ntt_mid_col_kernel<K>, K = A + B, A = ceil(K / 2)
inline int get_mid_tw1(DeviceTables &t, int dev, unsigned k, int inverse, const u64 **out, unsigned la_override = 0) {
    auto key = std::make_tuple(dev, k + 100 * la_override, inverse);
    auto it = g_mid_tw1.find(key);
    if (it != g_mid_tw1.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(k);
    if (inverse) w = hgl_inv(w);
    const u32 la = la_override ? la_override : (k + 1) / 2, lb = k - la;
    std::vector<u64> h((size_t)1 << k);
    for (u32 k1 = 0; k1 < (1u << la); k1++) {
        u64 step = hgl_pow(w, k1), acc = 1;
        for (u32 g = 0; g < (1u << lb); g++) {
            h[(k1 << lb) + g] = acc;
            acc = hgl_mul(acc, step);
        }
    }
    u64 *d;
    TF21_TRY(upload(t, h, &d));
    g_mid_tw1[key] = d;
    *out = d;
    return 0;
}

// [2^(K-5)][32] scalar * omega_{2^K}^(+-k1 b) for the 2^K-point kernel, 5 < K < 10
inline int get_small_n_tw(DeviceTables &t, int dev, unsigned k, int inverse, u64 scalar, const u64 **out) {
    auto key = std::make_tuple(dev, k, inverse, scalar);
    auto it = g_small_n_tw.find(key);
    if (it != g_small_n_tw.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(k);
    if (inverse) w = hgl_inv(w);
    const u32 rows = 1u << (k - 5);
    std::vector<u64> h(rows * 32);
    for (u32 k1 = 0; k1 < rows; k1++) {
        u64 step = hgl_pow(w, k1), acc = scalar % GL_P;
        for (u32 b = 0; b < 32; b++) {
            h[k1 * 32 + b] = acc;
            acc = hgl_mul(acc, step);
        }
    }
    u64 *d;
    TF21_TRY(upload(t, h, &d));
    g_small_n_tw[key] = d;
    *out = d;
    return 0;
}

inline int get_t1(DeviceTables &t, int dev, int inverse, const u64 **out) {
    FastTables &ft = g_fast_tables[dev];
    auto it = ft.t1.find(inverse);
    if (it != ft.t1.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(10);
    if (inverse) w = hgl_inv(w);
    std::vector<u64> h(1024);
    for (u32 k1 = 0; k1 < 32; k1++) {
        u64 step = hgl_pow(w, k1), acc = 1;
        for (u32 b = 0; b < 32; b++) {
            h[k1 * 32 + b] = acc;
            acc = hgl_mul(acc, step);
        }
    }
    u64 *d;
    TF21_TRY(upload(t, h, &d));
    ft.t1[inverse] = d;
    *out = d;
    return 0;
}

constexpr unsigned kFullTwiddleMaxLog = 20;  // 8 MiB per (size, direction): stays resident in L2

// T[j][i] = scalar * omega_B^(+-j i), j < B / N_p, i < N_p = 2^log_np
inline int get_tw_full(DeviceTables &t, int dev, unsigned log_b, unsigned log_np, int inverse, u64 scalar,
                       const u64 **out) {
    FastTables &ft = g_fast_tables[dev];
    auto key = std::make_tuple(log_b, log_np, inverse, scalar);
    auto it = ft.tw_full.find(key);
    if (it != ft.tw_full.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(log_b);
    if (inverse) w = hgl_inv(w);
    const u64 cols = 1ull << log_np;
    const u64 rows = (1ull << log_b) >> log_np;
    std::vector<u64> h(rows * cols);
    u64 step = 1;  // w^j
    for (u64 j = 0; j < rows; j++) {
        u64 acc = scalar % GL_P;
        for (u64 i = 0; i < cols; i++) {
            h[j * cols + i] = acc;
            acc = hgl_mul(acc, step);
        }
        step = hgl_mul(step, w);
    }
    u64 *d;
    TF21_TRY(upload(t, h, &d));
    ft.tw_full[key] = d;
    *out = d;
    return 0;
}

// small column passes: T[i][j] = scalar * omega_B^(+-i j), i < N_p = 2^log_np, j < B / N_p
inline int get_tw_small(DeviceTables &t, int dev, unsigned log_b, unsigned log_np, int inverse, u64 scalar,
                        const u64 **out) {
    FastTables &ft = g_fast_tables[dev];
    auto key = std::make_tuple(log_b, log_np, inverse, scalar);
    auto it = ft.tw_small.find(key);
    if (it != ft.tw_small.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(log_b);
    if (inverse) w = hgl_inv(w);
    const u64 rows = 1ull << log_np;
    const u64 cols = (1ull << log_b) >> log_np;
    std::vector<u64> h(rows * cols);
    u64 step = 1;  // w^i
    for (u64 i = 0; i < rows; i++) {
        u64 acc = scalar % GL_P;
        for (u64 j = 0; j < cols; j++) {
            h[i * cols + j] = acc;
            acc = hgl_mul(acc, step);
        }
        step = hgl_mul(step, w);
    }
    u64 *d;
    TF21_TRY(upload(t, h, &d));
    ft.tw_small[key] = d;
    *out = d;
    return 0;
}

// TF21_NO_TMA=1 in the environment keeps the LDGSTS-staged column pass (A/B runs, tools/ab.sh)
inline bool tma_disabled() {
    static const bool off = getenv("TF21_NO_TMA") != nullptr;
    return off;
}

// the unscale factor of an inverse transform is folded into the first twiddle table whenever there is a leading pass
inline bool post_scalar_is_foldable(u64 post_scalar, u32 n_lead) { return post_scalar == 0 || n_lead > 0; }
// Measured (2^20 x256): with the inter-pass twiddles applied on the load of the row pass the column pass drops
// 1.32 -> 1.08 ms but the row pass grows 1.10 -> 1.43 ms (the 32 extra products sit in front of the first butterflies,
// where all 32 elements are live) -- 2.52 ms against 2.42 ms, so the deferral stays off unless TF21_TW_DEFER is set.
inline bool tma_defer_disabled() {
    static const bool off = getenv("TF21_TW_DEFER") == nullptr;
    return off;
}
inline bool col_n_disabled() {
    static const bool off = getenv("TF21_NO_COL_N") != nullptr;
    return off;
}
inline bool col_n_forced() {
    static const bool on = getenv("TF21_COL_N_ALL") != nullptr;
    return on;
}
// TF21_MID_MASK: bit K set = the 2^K-point leading pass takes ntt_mid_col_kernel<K> (default: see kMidDefaultMask)
constexpr u32 kMidDefaultMask = TF21_MID_DEFAULT_MASK;
inline u32 mid_mask() {
    static const u32 m = [] {
        const char *e = getenv("TF21_MID_MASK");
        return e ? (u32)strtoul(e, nullptr, 0) : kMidDefaultMask;
    }();
    return m;
}
#ifndef TF21_MID2_DEFAULT_MASK
#define TF21_MID2_DEFAULT_MASK 0x280  /* bit K set: the 2^K-point leading pass (odd K) takes the two-thread form ntt_mid2_col_kernel: K = 9 (2^19: 1.63 -> 1.38 ms per GiB) and K = 7 (2^17: 1.211 -> 1.175, 2^27: 2.055 -> 2.005; profiles/r02l_ab_mid7_two_thread.txt) */
#endif
// TF21_MID2_MASK in the environment overrides the default for A/B runs
inline bool mid_two_thread(u32 k) {
    static const u32 m = [] {
        const char *e = getenv("TF21_MID2_MASK");
        return e ? (u32)strtoul(e, nullptr, 0) : (u32)TF21_MID2_DEFAULT_MASK;
    }();
    return (k == 5 || k == 7 || k == 9) && ((m >> k) & 1u);
}
inline bool small_n_disabled() {
    static const bool off = getenv("TF21_NO_SMALL_N") != nullptr;
    return off;
}
inline bool tma_row_disabled() {
    static const bool off = getenv("TF21_NO_TMA") != nullptr || getenv("TF21_NO_TMA_ROW") != nullptr;
    return off;
}

template <typename K, typename A>
inline int launch_fast_named(const char *name, K kernel, dim3 grid, const A &args, cudaStream_t st) {
    TF21_LAUNCH_NAMED(name, kernel, grid, kFastThreads, kFastSmem, st, args);
    return 0;
}
#define launch_fast(kernel, grid, args, st) launch_fast_named(#kernel, kernel, grid, args, st)

template <bool INV>
inline int launch_small(u32 a_log, unsigned grid, const SmallColArgs &args, cudaStream_t st) {
    switch (a_log) {
        case 1: TF21_LAUNCH_NAMED("ntt_small_col_kernel", (ntt_small_col_kernel<INV, 1>), grid, 128, 0, st, args); break;
        case 2: TF21_LAUNCH_NAMED("ntt_small_col_kernel", (ntt_small_col_kernel<INV, 2>), grid, 128, 0, st, args); break;
        case 3: TF21_LAUNCH_NAMED("ntt_small_col_kernel", (ntt_small_col_kernel<INV, 3>), grid, 128, 0, st, args); break;
        case 4: TF21_LAUNCH_NAMED("ntt_small_col_kernel", (ntt_small_col_kernel<INV, 4>), grid, 128, 0, st, args); break;
        case 5: TF21_LAUNCH_NAMED("ntt_small_col_kernel", (ntt_small_col_kernel<INV, 5>), grid, 128, 0, st, args); break;
        case 6: TF21_LAUNCH_NAMED("ntt_small_col_kernel", (ntt_small_col_kernel<INV, 6>), grid, 128, 0, st, args); break;
        default: return TF21_E_BAD_ARG;
    }
    return 0;
}

inline int launch_small_pruned(u32 a_log, u32 lnz, unsigned grid, const SmallColArgs &args, cudaStream_t st) {
#define TF21_PRUNED_CASE(A_, L_)                                                                          \
    if (a_log == (A_) && lnz == (L_)) {                                                                   \
        TF21_LAUNCH_NAMED("ntt_small_col_pruned_kernel", (ntt_small_col_pruned_kernel<A_, L_>), grid, 128, 0, st, \
                          args);                                                                          \
        return 0;                                                                                         \
    }
    TF21_PRUNED_CASE(1, 0)
    TF21_PRUNED_CASE(2, 0) TF21_PRUNED_CASE(2, 1)
    TF21_PRUNED_CASE(3, 0) TF21_PRUNED_CASE(3, 1) TF21_PRUNED_CASE(3, 2)
    TF21_PRUNED_CASE(4, 0) TF21_PRUNED_CASE(4, 1) TF21_PRUNED_CASE(4, 2) TF21_PRUNED_CASE(4, 3)
    TF21_PRUNED_CASE(5, 0) TF21_PRUNED_CASE(5, 1) TF21_PRUNED_CASE(5, 2) TF21_PRUNED_CASE(5, 3)
    TF21_PRUNED_CASE(6, 0) TF21_PRUNED_CASE(6, 1) TF21_PRUNED_CASE(6, 2) TF21_PRUNED_CASE(6, 3)
#undef TF21_PRUNED_CASE
    return TF21_E_BAD_ARG;
}

// number of leading rows of a 2^a_log-point first pass that can be non-zero, as a log2 (rounded up)
inline u32 nonzero_rows_log(u64 n_in, u64 inner_elems, u32 a_log) {
    u64 rows = (n_in + inner_elems - 1) / inner_elems;
    if (rows < 1) rows = 1;
    u32 l = 0;
    while (((u64)1 << l) < rows) l++;
    return l > a_log ? a_log : l;
}

// generic (shared-memory radix-2) path: sizes below 2^13 and width-3 single passes
inline int ntt_run_generic(DeviceTables &tabs, const u64 *src, u64 n_in, u64 *dst, u64 n, u32 w, u64 batch,
                           int inverse, ScaleTab pre, ScaleTab post, u64 post_scalar, u64 *scratch,
                           cudaStream_t st) {
    const u32 log_n = ilog2_u64(n);
    const NttPlan plan = make_plan(log_n);
    const u64 array_words = n * w;
    const u64 *tw_small = tabs.tw_small[inverse ? 1 : 0];
    const u64 *cur_src = src;
    u64 cur_src_words = n_in * w;
    u64 cur_n_in = n_in;
    ScaleTab cur_pre = pre;
    u32 consumed = 0;  // log2 of N_1..N_{p-1}
    for (u32 p = 0; p + 1 < plan.k; p++) {
        const u32 lp = plan.l[p];
        const u32 log_inner = log_n - consumed - lp;
        const u32 log_b = log_n - consumed;
        const u64 inner_words = ((u64)1 << log_inner) * w;
        const u32 n_outer = 1u << consumed;
        ColPassArgs a{};
        a.src = cur_src;
        a.dst = scratch;
        a.src_array_words = cur_src_words;
        a.dst_array_words = array_words;
        a.log_np = lp;
        a.w = w;
        a.inner_words = inner_words;
        a.n_col_tiles = (u32)((inner_words + kNttColTile - 1) / kNttColTile);
        a.n_outer = n_outer;
        a.n_in_elems = cur_n_in;
        a.tw_np = tw_small + ((1u << lp) >> 1) - 1;
        a.log_b = log_b;
        DeviceTables::Split sp;
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            TF21_TRY(get_split_tables(tabs, log_b, inverse, &sp));
        }
        a.tw = ScaleTab{sp.lo, sp.hi, sp.h};
        a.pre = cur_pre;
        u64 grid = batch * n_outer * a.n_col_tiles;
        if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
        u32 nt = pick_threads((u64)(1u << lp) / 2 * kNttColTile);
        TF21_LAUNCH(ntt_col_pass_kernel, (unsigned)grid, nt, col_pass_smem(lp), st, a);
        consumed += lp;
        cur_src = scratch;
        cur_src_words = array_words;
        cur_n_in = n;
        cur_pre = ScaleTab{nullptr, nullptr, 0};
    }

    const u32 lk = plan.l[plan.k - 1];
    const u32 nk = 1u << lk;
    RowPassArgs r{};
    r.src = cur_src;
    r.dst = dst;
    r.src_array_words = cur_src_words;
    r.dst_array_words = array_words;
    r.log_nk = lk;
    r.w = w;
    // rows per tile: keep the tile under ~150 KB of shared memory
    u32 to = 16;
    while (to > 1 && row_pass_smem(r.log_nk, to, w) > 150 * 1024) to--;
    r.tw_nk = tw_small + (nk >> 1) - 1;
    r.post = post;
    r.post_scalar = post_scalar;
    u64 grid;
    if (plan.k == 1) {
        r.single = 1;
        if (batch > 0xffffffffull) return TF21_E_LEN_TOO_LARGE;
        r.rows_total = (u32)batch;
        if ((u64)to > batch) to = (u32)batch;
        r.to = to;
        r.n_tiles_t = (u32)((batch + to - 1) / to);
        r.mid = 1;
        r.src_t_stride = cur_src_words;
        r.dst_t_stride = array_words;
        r.dst_i_stride = w;
        r.dst_mid_stride = 0;
        r.elem_i_stride = 1;
        r.elem_mid_stride = 0;
        r.n_in_elems = cur_n_in;
        r.pre = cur_pre;
        // the batch index is folded into the t axis: arrays are addressed through t strides
        r.src_array_words = 0;
        r.dst_array_words = 0;
        grid = r.n_tiles_t;
    } else {
        r.single = 0;
        const u32 n1 = 1u << plan.l[0];
        const u32 midc = (plan.k == 3) ? (1u << plan.l[1]) : 1u;
        if (to > n1) to = n1;
        while (n1 % to) to--;  // to must divide N_1 so tiles never straddle
        r.to = to;
        r.rows_total = n1;
        r.n_tiles_t = n1 / to;
        r.mid = midc;
        r.src_t_stride = (u64)midc * nk * w;  // consecutive i_1
        r.dst_t_stride = w;
        const u64 o_total = n >> r.log_nk;  // N_1 .. N_{k-1}
        r.dst_i_stride = o_total * w;
        r.dst_mid_stride = (u64)n1 * w;  // o' = i_1 + N_1 * i_2
        r.elem_i_stride = o_total;
        r.elem_mid_stride = n1;
        r.n_in_elems = nk;
        r.pre = ScaleTab{nullptr, nullptr, 0};
        grid = batch * r.n_tiles_t * r.mid;
    }
    if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
    u32 nt = pick_threads((u64)nk / 2 * r.to * w);
    TF21_LAUNCH(ntt_row_pass_kernel, (unsigned)grid, nt, row_pass_smem(r.log_nk, r.to, w), st, r);
    return 0;
}

// Core entry: dst[b] = scale_post( NTT_n( zero_extend( scale_pre( src[b][0..n_in) ) ) ) ) for b < batch.
// src arrays are n_in*w words apart, dst arrays n*w words apart.  src == dst is allowed when
// n_in == n.  `scratch` (n*w*batch words) is required when log2 n > 10.  Caller holds no lock;
// table lookups lock g_mutex internally.
//
// Pass plan for n >= 2^13 (register-resident kernels only):
//   [ one or two "small" column passes of 2^a <= 32 points ]  [ a 1024-point column pass if log2 n >= 20 ]
//   [ the 1024-point transposing row pass ]
// e.g. 2^20 = 1024 x 1024, 2^22 = 4 x 1024 x 1024, 2^26 = 8 x 8 x 1024 x 1024, 2^16 = 8 x 8 x 1024.
inline int ntt_run(DeviceTables &tabs, int dev, const u64 *src, u64 n_in, u64 *dst, u64 n, u32 w, u64 batch,
                   int inverse, ScaleTab pre, ScaleTab post, u64 post_scalar, u64 *scratch, cudaStream_t st) {
    const u32 log_n = ilog2_u64(n);
    const u64 array_words = n * w;
    if (log_n < 10) {
        // plain transforms (no zero extension, no coset scale tables) of whole columns: the register kernel
        if (n_in == n && !pre.lo && !post.lo && !small_n_disabled()) {
            SmallNArgs a{};
            a.src = src;
            a.dst = dst;
            a.total_elems = batch * n;
            a.w = w;
            if (log_n > 5) {
                std::lock_guard<std::mutex> lock(g_mutex);
                TF21_TRY(get_small_n_tw(tabs, dev, log_n, inverse, post_scalar ? post_scalar : 1, &a.tw));
                a.scaled = (post_scalar != 0 && post_scalar % GL_P != 1) ? 1u : 0u;
            } else {
                a.post_scalar = post_scalar;
            }
            const u64 chunks = (a.total_elems + 1023) / 1024;
            const u64 grid = (chunks * w + kFastCols - 1) / kFastCols;
            if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
#define TF21_SMALL_N_CASE(K_)                                                                                       \
    case K_:                                                                                                       \
        if (inverse)                                                                                               \
            TF21_LAUNCH_NAMED("ntt_small_n_kernel", (ntt_small_n_kernel<true, K_>), (unsigned)grid, kFastThreads,   \
                              kFastSmem, st, a);                                                                   \
        else                                                                                                       \
            TF21_LAUNCH_NAMED("ntt_small_n_kernel", (ntt_small_n_kernel<false, K_>), (unsigned)grid, kFastThreads,  \
                              kFastSmem, st, a);                                                                   \
        return 0;
            switch (log_n) {
                TF21_SMALL_N_CASE(1) TF21_SMALL_N_CASE(2) TF21_SMALL_N_CASE(3) TF21_SMALL_N_CASE(4) TF21_SMALL_N_CASE(5)
                TF21_SMALL_N_CASE(6) TF21_SMALL_N_CASE(7) TF21_SMALL_N_CASE(8) TF21_SMALL_N_CASE(9)
            }
#undef TF21_SMALL_N_CASE
        }
        return ntt_run_generic(tabs, src, n_in, dst, n, w, batch, inverse, pre, post, post_scalar, scratch, st);
    }

    const u64 *t1 = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        TF21_TRY(get_t1(tabs, dev, inverse, &t1));
    }
    if (log_n == 10) {
        FastSingleArgs a{};
        a.src = src;
        a.dst = dst;
        a.src_array_words = n_in * w;
        a.dst_array_words = array_words;
        a.batch = batch;
        a.n_in_elems = n_in;
        a.t1 = t1;
        a.post_scalar = post_scalar;
        a.pre = pre;
        a.post = post;
        u64 grid = (batch * w + kFastCols - 1) / kFastCols;
        if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
        if (w == 1) {
            if (inverse) return launch_fast_named("ntt1024_single_kernel", ntt1024_single_kernel<true, 1>, (unsigned)grid, a, st);
            return launch_fast_named("ntt1024_single_kernel", ntt1024_single_kernel<false, 1>, (unsigned)grid, a, st);
        }
        if (inverse) return launch_fast_named("ntt1024_single_kernel", ntt1024_single_kernel<true, 3>, (unsigned)grid, a, st);
        return launch_fast_named("ntt1024_single_kernel", ntt1024_single_kernel<false, 3>, (unsigned)grid, a, st);
    }

    // ---- plan: logs of the leading (column) passes, the last pass is always the 1024-point row pass ----
    u32 lead[3];
    u32 n_lead = 0;
    bool lead_is_col10[3] = {false, false, false};
    bool first_is_coln = false, first_is_mid = false;
    {
        u32 rem = log_n - 10;
        const bool has_col = rem >= 10;
        if (has_col) rem -= 10;
        // zero-extended input (coset evaluate / LDE): one pruned pass over all the small bits reads only the
        // non-zero rows and replaces two full passes
        const bool prune_all = !inverse && n_in < n && rem >= 1 && rem <= 6 &&
                               nonzero_rows_log(n_in, n >> rem, rem) <= 3 && nonzero_rows_log(n_in, n >> rem, rem) < rem;
        // The register pass (ntt_col_n_kernel) is taken where it measured faster than the thread-per-column passes
        // (profiles/r02e_ntt_size_sweep.txt): 2- and 4-point leading passes (2^11, 2^12, 2^21, 2^22: -6 .. -19 %) and
        // the 64-point pass in front of a 1024-point column pass (2^26: one pass fewer).  In between it is latency
        // bound (4 warps per scheduler, 41 % issue, profiles/r02e_ncu_col_n_summary.txt) and loses 1 .. 15 %.
        first_is_coln = !prune_all && n_in == n && !pre.lo && !col_n_disabled() &&
                        ((rem >= 1 && rem <= 2) || (rem == 6 && has_col) || (col_n_forced() && rem >= 1 && rem <= 9));
        // 32 .. 512 points in one pass through shared memory (ntt_mid_col_kernel) instead of two thread-per-column passes
        first_is_mid = !prune_all && n_in == n && !pre.lo && rem >= 3 && rem <= 9 && ((mid_mask() >> rem) & 1u) &&
                       !(col_n_forced() && first_is_coln) && ((((u64)1 << (log_n - rem)) * w) % (kMidThreads >> (rem / 2))) == 0 && (rem != 9 || mid_two_thread(9));
        if (first_is_mid) first_is_coln = true;  // same table and argument set-up as the register pass
        if (prune_all || first_is_coln) {
            lead[n_lead++] = rem;  // one pruned pass, or one register pass of 2^rem points (ntt_col_n_kernel)
        } else if (rem > 5) {  // 64-point columns per thread do not pay: ~200 KB of straight-line code, 196 registers
            lead[n_lead++] = (rem + 1) / 2;
            lead[n_lead++] = rem / 2;
        } else if (rem > 0) {
            lead[n_lead++] = rem;
        }
        if (has_col) {
            lead_is_col10[n_lead] = true;
            lead[n_lead++] = 10;
        }
    }

    const u64 *cur_src = src;
    u64 cur_src_words = n_in * w;
    u64 cur_n_in = n_in;
    ScaleTab cur_pre = pre;
    u32 consumed = 0;
    // will the last pass take the TMA-store form?  (then it can also apply the inter-pass twiddles of a TMA column pass)
    const bool row_tma_ok = post_scalar_is_foldable(post_scalar, n_lead) && post.lo == nullptr && !tma_row_disabled() &&
                            ((n >> 10) * w) % kFastCols == 0 && array_words < (1ull << 29) && ((uintptr_t)dst & 15) == 0 &&
                            tma_encoder() != nullptr;
    const u64 *deferred_tw = nullptr;
    for (u32 p = 0; p < n_lead; p++) {
        const u32 lp = lead[p];
        const u32 log_inner = log_n - consumed - lp;
        const u32 log_b = log_n - consumed;
        const u64 inner_words = ((u64)1 << log_inner) * w;
        const u32 n_outer = 1u << consumed;
        const bool is_coln = p == 0 && first_is_coln;
        u64 scalar = 1;
        // fold the unscale (ntt.rs:220-228) into the first twiddle TABLE: a register pass that works from the split
        // tables (B > 2^20) leaves it to the 1024-point column pass behind it, whose table always exists
        if (post_scalar && !(is_coln && log_b > kFullTwiddleMaxLog)) {
            scalar = post_scalar;
            post_scalar = 0;
        }
        const u64 *tw_full = nullptr;
        ScaleTab tw{nullptr, nullptr, 0};
        u64 tw_scalar = 0;
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            if (log_b <= kFullTwiddleMaxLog) {
                if (lead_is_col10[p])
                    TF21_TRY(get_tw_full(tabs, dev, log_b, lp, inverse, scalar, &tw_full));
                else
                    TF21_TRY(get_tw_small(tabs, dev, log_b, lp, inverse, scalar, &tw_full));
            } else {
                DeviceTables::Split sp;
                TF21_TRY(get_split_tables(tabs, log_b, inverse, &sp));
                tw = ScaleTab{sp.lo, sp.hi, sp.h};
                tw_scalar = scalar == 1 ? 0 : scalar;
            }
        }
        if (is_coln) {
            ColNArgs a{};
            a.src = cur_src;
            a.dst = scratch;
            a.array_words = array_words;
            a.inner_words = (u32)inner_words;
            a.w = w;
            a.n_outer = n_outer;
            a.tw_full = tw_full;
            a.tw = tw;
            a.log_b = log_b;
            a.scaled = scalar % GL_P != 1 ? 1u : 0u;
            if (first_is_mid) {
                {
                    std::lock_guard<std::mutex> lock(g_mutex);
                    TF21_TRY(get_mid_tw1(tabs, dev, lp, inverse, &a.tw1, mid_two_thread(lp) ? (lp - 1) / 2 : 0u));
                }
                // with a full table row 0 carries the scalar (or ones): every row is multiplied, nothing to flag
                for (u64 b0 = 0; b0 < batch; b0 += 65535) {
                    a.src = cur_src + b0 * array_words;
                    a.dst = scratch + b0 * array_words;
                    const unsigned nb = (unsigned)(batch - b0 < 65535 ? batch - b0 : 65535);
#define TF21_MID_CASE(K_)                                                                                             \
    case K_: {                                                                                                       \
        const dim3 grid((unsigned)(inner_words / MidShape<K_>::C), n_outer, nb);                                     \
        static std::once_flag once_f, once_i;                                                                        \
        if (inverse) {                                                                                               \
            std::call_once(once_i, [] {                                                                              \
                cudaFuncSetAttribute(ntt_mid_col_kernel<true, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)MidShape<K_>::smem);                                                       \
            });                                                                                                      \
            TF21_LAUNCH_NAMED("ntt_mid_col_kernel", (ntt_mid_col_kernel<true, K_>), grid, kMidThreads,                \
                              MidShape<K_>::smem, st, a);                                                            \
        } else {                                                                                                     \
            std::call_once(once_f, [] {                                                                              \
                cudaFuncSetAttribute(ntt_mid_col_kernel<false, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)MidShape<K_>::smem);                                                       \
            });                                                                                                      \
            TF21_LAUNCH_NAMED("ntt_mid_col_kernel", (ntt_mid_col_kernel<false, K_>), grid, kMidThreads,               \
                              MidShape<K_>::smem, st, a);                                                            \
        }                                                                                                            \
        break;                                                                                                       \
    }
#define TF21_MID2_CASE(HB_)                                                                                           \
    {                                                                                                                \
        const dim3 grid((unsigned)(inner_words / Mid2Shape<HB_>::C), n_outer, nb);                                   \
        if (inverse)                                                                                                 \
            TF21_LAUNCH_NAMED("ntt_mid2_col_kernel", (ntt_mid2_col_kernel<true, HB_>), grid, kMidThreads,             \
                              Mid2Shape<HB_>::smem, st, a);                                                          \
        else                                                                                                         \
            TF21_LAUNCH_NAMED("ntt_mid2_col_kernel", (ntt_mid2_col_kernel<false, HB_>), grid, kMidThreads,            \
                              Mid2Shape<HB_>::smem, st, a);                                                          \
        continue;                                                                                                    \
    }
                    if (mid_two_thread(lp)) {
                        if (lp == 5) TF21_MID2_CASE(2)
                        if (lp == 7) TF21_MID2_CASE(3)
                        TF21_MID2_CASE(4)
                    }
#undef TF21_MID2_CASE
                    switch (lp) {
                        TF21_MID_CASE(3) TF21_MID_CASE(4) TF21_MID_CASE(5) TF21_MID_CASE(6) TF21_MID_CASE(7) TF21_MID_CASE(8)
                        TF21_MID_CASE(9)
                        default: return TF21_E_BAD_ARG;
                    }
#undef TF21_MID_CASE
                }
                consumed += lp;
                cur_src = scratch;
                cur_src_words = array_words;
                cur_n_in = n;
                cur_pre = ScaleTab{nullptr, nullptr, 0};
                continue;
            }
            if (lp > 5) {
                std::lock_guard<std::mutex> lock(g_mutex);
                TF21_TRY(get_col_n_tw1(tabs, dev, lp, inverse, &a.tw1));
            }
            const u32 groups = (u32)(inner_words >> (10 - lp));  // word-column groups of G = 2^(10-lp) per block of rows
            for (u64 b0 = 0; b0 < batch; b0 += 65535) {
                a.src = cur_src + b0 * array_words;
                a.dst = scratch + b0 * array_words;
                const dim3 grid((groups + kFastCols - 1) / kFastCols, n_outer, (unsigned)(batch - b0 < 65535 ? batch - b0 : 65535));
#define TF21_COL_N_CASE(K_)                                                                                          \
    case K_:                                                                                                        \
        if (inverse)                                                                                                \
            TF21_LAUNCH_NAMED("ntt_col_n_kernel", (ntt_col_n_kernel<true, K_>), grid, kFastThreads,                 \
                              (K_ <= 5 ? 0 : kFastSmem), st, a);                                                    \
        else                                                                                                        \
            TF21_LAUNCH_NAMED("ntt_col_n_kernel", (ntt_col_n_kernel<false, K_>), grid, kFastThreads,                \
                              (K_ <= 5 ? 0 : kFastSmem), st, a);                                                    \
        break;
                switch (lp) {
                    TF21_COL_N_CASE(1) TF21_COL_N_CASE(2) TF21_COL_N_CASE(3) TF21_COL_N_CASE(4) TF21_COL_N_CASE(5)
                    TF21_COL_N_CASE(6) TF21_COL_N_CASE(7) TF21_COL_N_CASE(8) TF21_COL_N_CASE(9)
                    default: return TF21_E_BAD_ARG;
                }
#undef TF21_COL_N_CASE
            }
        } else if (lead_is_col10[p]) {
            FastColArgs a{};
            a.src = cur_src;
            a.dst = scratch;
            a.src_array_words = cur_src_words;
            a.dst_array_words = array_words;
            a.w = w;
            a.inner_words = inner_words;
            a.n_col_tiles = (u32)(inner_words / kFastCols);
            a.n_outer = n_outer;
            a.n_in_elems = cur_n_in;
            a.t1 = t1;
            a.tw_full = tw_full;
            a.tw = tw;
            a.log_b = log_b;
            a.tw_scalar = tw_scalar;
            a.pre = cur_pre;
            if (n_outer > 65535 || inner_words > (1u << 21)) return TF21_E_LEN_TOO_LARGE;
            const bool plain = cur_n_in == n && !cur_pre.lo && tw_full != nullptr;
            // PLAIN column passes move their tiles by TMA: source and scratch are [batch * n_outer][1024][inner_words]
            // tensors (source arrays are n * w words apart when every input row exists)
            if (plain && !tma_disabled()) {
                CUtensorMap src_map, dst_map;
                const u64 slabs = batch * n_outer;
                if (tma_encode_tile_map(&src_map, cur_src, inner_words, slabs) &&
                    tma_encode_tile_map(&dst_map, scratch, inner_words, slabs)) {
                    ColTmaArgs ta{t1, tw_full, w, 0};
                    const bool defer = row_tma_ok && log_b == 20 && !tma_defer_disabled();
                    if (defer) deferred_tw = tw_full;
                    for (u64 s0 = 0; s0 < slabs; s0 += 65535) {
                        ta.slab0 = (u32)s0;
                        const dim3 grid((unsigned)(inner_words / kTmaTileCols),
                                        (unsigned)(slabs - s0 < 65535 ? slabs - s0 : 65535));
#define TF21_COL_TMA_LAUNCH(I_, T_)                                                                                      \
    TF21_LAUNCH_NAMED("ntt1024_col_tma_kernel<" #I_ ">", (ntt1024_col_tma_kernel<I_, T_>), grid, kFastThreads, kColTmaSmem, \
                      st, src_map, dst_map, ta)
                        if (inverse) {
                            if (defer) TF21_COL_TMA_LAUNCH(true, false);
                            else TF21_COL_TMA_LAUNCH(true, true);
                        } else {
                            if (defer) TF21_COL_TMA_LAUNCH(false, false);
                            else TF21_COL_TMA_LAUNCH(false, true);
                        }
#undef TF21_COL_TMA_LAUNCH
                    }
                    consumed += lp;
                    cur_src = scratch;
                    cur_src_words = array_words;
                    cur_n_in = n;
                    cur_pre = ScaleTab{nullptr, nullptr, 0};
                    continue;
                }
            }
            // the batch index is the z dimension of the grid (<= 65535): larger batches go in slices
            for (u64 b0 = 0; b0 < batch; b0 += 65535) {
                const u64 nb = batch - b0 < 65535 ? batch - b0 : 65535;
                a.src = cur_src + b0 * cur_src_words;
                a.dst = scratch + b0 * array_words;
                const dim3 grid(a.n_col_tiles, n_outer, (unsigned)nb);
                if (inverse) {
                    if (plain) TF21_TRY(launch_fast_named("ntt1024_col_kernel<true>", (ntt1024_col_kernel<true, true>), grid, a, st));
                    else TF21_TRY(launch_fast_named("ntt1024_col_kernel<true>", (ntt1024_col_kernel<true, false>), grid, a, st));
                } else {
                    if (plain) TF21_TRY(launch_fast_named("ntt1024_col_kernel<false>", (ntt1024_col_kernel<false, true>), grid, a, st));
                    else TF21_TRY(launch_fast_named("ntt1024_col_kernel<false>", (ntt1024_col_kernel<false, false>), grid, a, st));
                }
            }
        } else {
            SmallColArgs a{};
            a.src = cur_src;
            a.dst = scratch;
            a.src_array_words = cur_src_words;
            a.dst_array_words = array_words;
            a.w = w;
            a.inner_words = inner_words;
            a.n_outer = n_outer;
            a.n_in_elems = cur_n_in;
            a.tw_full = tw_full;
            a.tw = tw;
            a.tw_scalar = tw_scalar;
            a.pre = cur_pre;
            u64 grid = batch * n_outer * (inner_words / 128);
            if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
            const u32 lnz = (p == 0 && !inverse) ? nonzero_rows_log(cur_n_in, inner_words / w, lp) : lp;
            if (lnz < lp && lnz <= 3)
                TF21_TRY(launch_small_pruned(lp, lnz, (unsigned)grid, a, st));
            else if (inverse)
                TF21_TRY(launch_small<true>(lp, (unsigned)grid, a, st));
            else
                TF21_TRY(launch_small<false>(lp, (unsigned)grid, a, st));
        }
        consumed += lp;
        cur_src = scratch;
        cur_src_words = array_words;
        cur_n_in = n;
        cur_pre = ScaleTab{nullptr, nullptr, 0};
    }

    FastRowArgs a{};
    a.src = cur_src;
    a.dst = dst;
    a.array_words = array_words;
    a.w = w;
    a.n1 = 1u << lead[0];
    a.n2 = n_lead > 1 ? 1u << lead[1] : 1u;
    a.n3 = n_lead > 2 ? 1u << lead[2] : 1u;
    a.log_n1 = lead[0];
    a.log_n2 = n_lead > 1 ? lead[1] : 0u;
    const u64 rows = n >> 10;
    a.n_cols_total = batch * rows * w;
    a.n_tiles = (rows * w) % kFastCols == 0 ? (u32)(rows * w / kFastCols) : 0u;
    a.t1 = t1;
    a.post_scalar = post_scalar;
    a.post = post;
    // 16-byte stores need a 16-byte aligned destination for every array of the batch; a view at an odd word offset
    // (8-byte aligned only) takes the 8-byte path
    a.store128 = (((uintptr_t)dst & 15) == 0 && (array_words & 1) == 0) ? 1u : 0u;
    dim3 grid;
    if (a.n_tiles && batch <= 65535) {
        grid = dim3(a.n_tiles, (unsigned)batch);
    } else {
        a.n_tiles = 0;
        const u64 g = (a.n_cols_total + kFastCols - 1) / kFastCols;
        if (g > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
        grid = dim3((unsigned)g);
    }
    const bool has_post = post_scalar != 0 || post.lo != nullptr;
    // tiled shapes without a post-scale and with an aligned destination send their output tiles by TMA
    if (!has_post && a.n_tiles && !tma_row_disabled() && array_words < (1ull << 29)) {
        CUtensorMap dst_map;
        if (tma_encode_tile_map(&dst_map, dst, rows * w, batch)) {
            RowTmaArgs ra{cur_src, array_words, a.n1, a.n2, a.n3, a.log_n1, a.log_n2, t1, 0, deferred_tw};
            for (u64 b0 = 0; b0 < batch; b0 += 65535) {
                ra.b0 = (u32)b0;
                const dim3 g2(a.n_tiles, (unsigned)(batch - b0 < 65535 ? batch - b0 : 65535));
#define TF21_ROW_TMA_LAUNCH(I_, W_)                                                                                      \
    do {                                                                                                                \
        if (deferred_tw)                                                                                                \
            TF21_LAUNCH_NAMED("ntt1024_row_tma_kernel", (ntt1024_row_tma_kernel<I_, W_, true>), g2, kFastThreads,       \
                              kColTmaSmem, st, dst_map, ra);                                                            \
        else                                                                                                            \
            TF21_LAUNCH_NAMED("ntt1024_row_tma_kernel", (ntt1024_row_tma_kernel<I_, W_, false>), g2, kFastThreads,      \
                              kColTmaSmem, st, dst_map, ra);                                                            \
    } while (0)
                if (w == 1) {
                    if (inverse) TF21_ROW_TMA_LAUNCH(true, 1);
                    else TF21_ROW_TMA_LAUNCH(false, 1);
                } else {
                    if (inverse) TF21_ROW_TMA_LAUNCH(true, 3);
                    else TF21_ROW_TMA_LAUNCH(false, 3);
                }
#undef TF21_ROW_TMA_LAUNCH
            }
            return 0;
        }
    }
    if (deferred_tw) {  // unreachable by construction of row_tma_ok; never return an untwiddled transform
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "internal: deferred twiddles without a TMA row pass");
        return TF21_E_CUDA;
    }
#define TF21_ROW_LAUNCH(I_, W_, P_) \
    return launch_fast_named("ntt1024_row_kernel", (ntt1024_row_kernel<I_, W_, P_>), grid, a, st)
    if (w == 1) {
        if (inverse) { if (has_post) TF21_ROW_LAUNCH(true, 1, true); TF21_ROW_LAUNCH(true, 1, false); }
        if (has_post) TF21_ROW_LAUNCH(false, 1, true);
        TF21_ROW_LAUNCH(false, 1, false);
    }
    if (inverse) { if (has_post) TF21_ROW_LAUNCH(true, 3, true); TF21_ROW_LAUNCH(true, 3, false); }
    if (has_post) TF21_ROW_LAUNCH(false, 3, true);
    TF21_ROW_LAUNCH(false, 3, false);
#undef TF21_ROW_LAUNCH
}

}  // namespace tf21
