// ntt_kernels.cuh -- batched Goldilocks NTT / iNTT (K1, K2 of SURVEY.md), generic multi-pass path.
//
// Replaces math::ntt::{ntt, intt, ntt_unchecked, unscale} (twenty-first/src/math/ntt.rs:67-82,
// 109-125, 153-228) and the scale / zero-pad steps of Polynomial::fast_coset_{evaluate,interpolate}
// (math/polynomial.rs:760-773, 1374-1399, 1907-1918).
//
// Functional spec: for every array of the batch and every coefficient lane
// c < w, y[i][c] = sum_j x[j][c] * omega_n^(i j) mod p on the raw words, natural order in and out,
// omega_n = 7^((p-1)/n) (PRIMITIVE_ROOTS, b_field_element.rs:43-78).  The reference does a
// bit-reversal pass plus log2(n) radix-2 DIT passes over memory; here the transform is split into
// k = ceil(log2 n / 10) passes of at most 1024 points each (digit decomposition of the index):
//
//   input index  j = j_1 (N_2..N_k) + j_2 (N_3..N_k) + ... + j_k
//   output index i = i_1 + N_1 i_2 + N_1 N_2 i_3 + ...
//
//   column pass p < k : position preserving; view [outer][N_p][inner = N_{p+1}..N_k][w], DFT along
//                       N_p for a tile of 16 consecutive words of the inner axis (128 B segments),
//                       then multiply by omega_{N_p inner}^(i_p * j_rest).
//   row pass (last)   : DFT along the contiguous N_k axis for TO rows that differ in i_1, stored
//                       transposed so that consecutive i_1 are consecutive in memory.
//
// Inside a CTA the DFT is an in-shared-memory radix-2 DIF (natural in, bit-reversed out; the
// un-reversal is folded into the store).  All arithmetic is canonical mod p.
//
// HBM layout: arrays are dense, n*w words each, AoS for w = 3 (x_field_element.rs:56-59).  A
// scratch buffer of the same size carries the data between passes.
#pragma once
#include "runtime.cuh"

namespace tf21 {

constexpr u32 kNttMaxLogPass = 10;  // at most 1024 points per pass
constexpr u32 kNttColTile = 16;     // words per row segment in a column pass (128 B)
constexpr u32 kNttColStride = 17;   // padded shared-memory row stride (odd => conflict free)

struct ScaleTab {  // factor(idx) = lo[idx & (2^h - 1)] * hi[idx >> h]; lo == nullptr => none
    const u64 *lo;
    const u64 *hi;
    u32 h;
};

__device__ __forceinline__ u64 scale_factor(const ScaleTab &t, u64 idx) {
    return gl_mulc(t.lo[idx & ((1ull << t.h) - 1)], t.hi[idx >> t.h]);
}

// shared-memory row padding: one spare row every 32 so that bit-reversed row reads spread over banks
__device__ __forceinline__ u32 rowp(u32 r) { return r + (r >> 5); }

__device__ __forceinline__ u32 bitrev(u32 v, u32 bits) { return __brev(v) >> (32 - bits); }

// radix-2 DIF over the rows of tile[rowp(r) * stride + c], c < ncol.  Natural in, bit-reversed out.
// tw[e] = omega_{np}^e, e < np/2.   (butterfly of ntt.rs:203-210 in decimation-in-frequency form)
__device__ __forceinline__ void dft_dif_smem(u64 *tile, u32 stride, u32 ncol, u32 log_np, const u64 *tw) {
    const u32 items = (1u << (log_np - 1)) * ncol;
    for (u32 s = 0; s < log_np; s++) {
        const u32 lh = log_np - 1 - s;
        for (u32 it = threadIdx.x; it < items; it += blockDim.x) {
            u32 c = it % ncol, bf = it / ncol;
            u32 pos = bf & ((1u << lh) - 1), grp = bf >> lh;
            u32 r0 = (grp << (lh + 1)) + pos, r1 = r0 + (1u << lh);
            u64 *p0 = tile + rowp(r0) * stride + c;
            u64 *p1 = tile + rowp(r1) * stride + c;
            u64 u = *p0, v = *p1;
            *p0 = gl_add(u, v);
            *p1 = gl_mulc(gl_sub(u, v), tw[pos << s]);
        }
        __syncthreads();
    }
}

struct ColPassArgs {
    const u64 *src;
    u64 *dst;
    u64 src_array_words, dst_array_words;
    u32 log_np, w;
    u64 inner_words;  // inner elements * w = row stride in words
    u32 n_col_tiles, n_outer;
    u64 n_in_elems;   // elements with index >= n_in_elems read as zero (zero extension, pass 1 only)
    const u64 *tw_np;
    ScaleTab tw;      // omega_B^e, B = N_p * inner
    u32 log_b;
    ScaleTab pre;     // optional per-element scale on load, indexed by the element index in the array
};

__global__ void ntt_col_pass_kernel(const ColPassArgs a) {
    extern __shared__ u64 smem[];
    const u32 np = 1u << a.log_np;
    u64 *tile = smem;
    u64 *twsm = smem + (size_t)(np + (np >> 5) + 1) * kNttColStride;
    const u32 ct = blockIdx.x % a.n_col_tiles;
    const u32 rest = blockIdx.x / a.n_col_tiles;
    const u32 o = rest % a.n_outer, b = rest / a.n_outer;
    const u64 q0 = (u64)ct * kNttColTile;
    const u32 ncols = (u32)min((u64)kNttColTile, a.inner_words - q0);
    const u64 inner_elems = a.inner_words / a.w;

    for (u32 i = threadIdx.x; i < (np >> 1); i += blockDim.x) twsm[i] = a.tw_np[i];

    const u64 block_off = (u64)o * np * a.inner_words + q0;
    const u64 *src = a.src + (u64)b * a.src_array_words + block_off;
    for (u32 idx = threadIdx.x; idx < np * kNttColTile; idx += blockDim.x) {
        u32 r = idx / kNttColTile, c = idx % kNttColTile;
        u64 v = 0;
        if (c < ncols) {
            u64 j = ((u64)o * np + r) * inner_elems + (q0 + c) / a.w;
            if (j < a.n_in_elems) {
                v = gl_canon(src[(u64)r * a.inner_words + c]);
                if (a.pre.lo) v = gl_mulc(v, scale_factor(a.pre, j));
            }
        }
        tile[rowp(r) * kNttColStride + c] = v;
    }
    __syncthreads();

    dft_dif_smem(tile, kNttColStride, kNttColTile, a.log_np, twsm);

    u64 *dst = a.dst + (u64)b * a.dst_array_words + block_off;
    const u64 bmask = (1ull << a.log_b) - 1;
    for (u32 idx = threadIdx.x; idx < np * kNttColTile; idx += blockDim.x) {
        u32 ip = idx / kNttColTile, c = idx % kNttColTile;
        if (c < ncols) {
            u64 v = tile[rowp(bitrev(ip, a.log_np)) * kNttColStride + c];
            u64 jrest = (q0 + c) / a.w;
            u64 e = ((u64)ip * jrest) & bmask;
            dst[(u64)ip * a.inner_words + c] = gl_mulc(v, scale_factor(a.tw, e));
        }
    }
}

struct RowPassArgs {
    const u64 *src;
    u64 *dst;
    u64 src_array_words, dst_array_words;
    u32 log_nk, w, to;       // to = rows per tile
    u32 n_tiles_t, mid;      // tiles along the t axis; number of `mid` values (1 unless 3 passes)
    u32 rows_total;          // N_1 (multi pass) or batch (single pass)
    u32 single;              // 1: single-pass mode (rows are whole arrays of the batch)
    u64 src_t_stride;        // words between consecutive rows of a tile in src
    u64 dst_t_stride, dst_i_stride, dst_mid_stride;
    u64 elem_i_stride, elem_mid_stride;  // output element index = t (multi) + mid*.. + i_k * elem_i_stride
    u64 n_in_elems;          // single-pass zero extension
    const u64 *tw_nk;
    ScaleTab pre;            // single-pass only (indexed by input element index)
    ScaleTab post;           // indexed by output element index
    u64 post_scalar;         // 0 => none ; else multiply every output by it (n^-1 for intt)
};

__global__ void ntt_row_pass_kernel(const RowPassArgs a) {
    extern __shared__ u64 smem[];
    const u32 nk = 1u << a.log_nk;
    const u32 ncol = a.to * a.w;
    const u32 stride = ncol | 1;
    u64 *tile = smem;
    u64 *twsm = smem + (size_t)(nk + (nk >> 5) + 1) * stride;
    const u32 tt = blockIdx.x % a.n_tiles_t;
    const u32 rest = blockIdx.x / a.n_tiles_t;
    const u32 mid = rest % a.mid, b = rest / a.mid;
    const u32 t0 = tt * a.to;
    const u32 rows_valid = min(a.to, a.rows_total - t0);
    const u32 row_words = nk * a.w;

    for (u32 i = threadIdx.x; i < (nk >> 1); i += blockDim.x) twsm[i] = a.tw_nk[i];

    const u64 *src = a.src + (u64)b * a.src_array_words + (u64)t0 * a.src_t_stride + (u64)mid * row_words;
    for (u32 idx = threadIdx.x; idx < a.to * row_words; idx += blockDim.x) {
        u32 t = idx / row_words, rw = idx % row_words;
        u32 r = rw / a.w, lane = rw % a.w;
        u64 v = 0;
        if (t < rows_valid && (u64)r < a.n_in_elems) {
            v = gl_canon(src[(u64)t * a.src_t_stride + rw]);
            if (a.pre.lo) v = gl_mulc(v, scale_factor(a.pre, r));
        }
        tile[rowp(r) * stride + t * a.w + lane] = v;
    }
    __syncthreads();

    dft_dif_smem(tile, stride, ncol, a.log_nk, twsm);

    u64 *dst = a.dst + (u64)b * a.dst_array_words + (u64)t0 * a.dst_t_stride + (u64)mid * a.dst_mid_stride;
    const u64 elem_base = (a.single ? 0 : (u64)t0) + (u64)mid * a.elem_mid_stride;
    if (!a.single) {
        // consecutive (t, lane) are consecutive in memory
        for (u32 idx = threadIdx.x; idx < nk * ncol; idx += blockDim.x) {
            u32 ik = idx / ncol, cc = idx % ncol;
            u32 t = cc / a.w;
            if (t < rows_valid) {
                u64 v = tile[rowp(bitrev(ik, a.log_nk)) * stride + cc];
                if (a.post_scalar) v = gl_mulc(v, a.post_scalar);
                if (a.post.lo) v = gl_mulc(v, scale_factor(a.post, elem_base + t + (u64)ik * a.elem_i_stride));
                dst[(u64)ik * a.dst_i_stride + cc] = v;
            }
        }
    } else {
        // consecutive (i_k, lane) are consecutive in memory
        for (u32 idx = threadIdx.x; idx < a.to * row_words; idx += blockDim.x) {
            u32 t = idx / row_words, rw = idx % row_words;
            u32 ik = rw / a.w, lane = rw % a.w;
            if (t < rows_valid) {
                u64 v = tile[rowp(bitrev(ik, a.log_nk)) * stride + t * a.w + lane];
                if (a.post_scalar) v = gl_mulc(v, a.post_scalar);
                if (a.post.lo) v = gl_mulc(v, scale_factor(a.post, (u64)ik));
                dst[(u64)t * a.dst_t_stride + rw] = v;
            }
        }
    }
}

// ---- Hadamard product for Polynomial::fast_multiply (polynomial.rs:923-927) ------------------------
// The operands are raw Montgomery words, so the reference's `l * r` is (l_raw * r_raw) * R^-1 mod p per
// BFieldElement product; every term of the XFieldElement product (x_field_element.rs:512-535) is a
// product of exactly two words, so the whole result is the plain mod-p product times `rinv` = 2^-64.
// Output words are lazy (any u64): the inverse transform that follows accepts any representative.
// (a and b may be the same buffer: Polynomial::fast_square, polynomial.rs:780-802; every thread reads its
// element of both operands before it writes)
__global__ void hadamard_kernel(u64 *a, const u64 *b, u64 n_elems, u32 w, u64 rinv) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    if (w == 1) {
        a[i] = gl_mul(gl_mul(a[i], b[i]), rinv);
        return;
    }
    const u64 c = a[3 * i], bb = a[3 * i + 1], aa = a[3 * i + 2];
    const u64 f = b[3 * i], e = b[3 * i + 1], d = b[3 * i + 2];
    const u64 ae = gl_mulc(aa, e), bd = gl_mulc(bb, d), ad = gl_mulc(aa, d);
    const u64 r0 = gl_sub(gl_sub(gl_mulc(c, f), ae), bd);
    const u64 r1 = gl_add(gl_add(gl_sub(gl_add(gl_mulc(bb, f), gl_mulc(c, e)), ad), ae), bd);
    const u64 r2 = gl_add(gl_add(gl_add(gl_mulc(aa, f), gl_mulc(bb, e)), gl_mulc(c, d)), ad);
    a[3 * i] = gl_mul(r0, rinv);
    a[3 * i + 1] = gl_mul(r1, rinv);
    a[3 * i + 2] = gl_mul(r2, rinv);
}

// ---- out-of-domain evaluation (next wave) --------------------------------------------
// Polynomial::evaluate (Horner, polynomial.rs:309-319) for many (polynomial, point) pairs: the tail of
// par_batch_coset_extrapolate (polynomial.rs:2255-2331).  Coefficients and accumulators are raw Montgomery
// words; the point is passed as its canonical *values* x' = raw * 2^-64, so that the plain mod-p product
// acc * x' is the raw word of acc * x (for XFieldElements term by term, x_field_element.rs:512-535).
template <u32 W>
struct EvalElem {
    u64 c[W];
};
template <u32 W>
__device__ __forceinline__ EvalElem<W> eval_mul(const EvalElem<W> &l, const EvalElem<W> &r) {
    EvalElem<W> o;
    if constexpr (W == 1) {
        o.c[0] = gl_mulc(l.c[0], r.c[0]);
    } else {
        const u64 c = l.c[0], b = l.c[1], a = l.c[2], f = r.c[0], e = r.c[1], d = r.c[2];
        const u64 ae = gl_mulc(a, e), bd = gl_mulc(b, d), ad = gl_mulc(a, d);
        o.c[0] = gl_sub(gl_sub(gl_mulc(c, f), ae), bd);
        o.c[1] = gl_add(gl_add(gl_sub(gl_add(gl_mulc(b, f), gl_mulc(c, e)), ad), ae), bd);
        o.c[2] = gl_add(gl_add(gl_add(gl_mulc(a, f), gl_mulc(b, e)), gl_mulc(c, d)), ad);
    }
    return o;
}
template <u32 W>
__device__ __forceinline__ EvalElem<W> eval_add(const EvalElem<W> &l, const EvalElem<W> &r) {
    EvalElem<W> o;
#pragma unroll
    for (u32 k = 0; k < W; k++) o.c[k] = gl_add(l.c[k], r.c[k]);
    return o;
}

constexpr u32 kEvalThreads = 256;
// grid = (n_points, n_polys).  polys: n_polys arrays of n elements (canonical raw words); points_val: canonical
// values of the points; out[(poly * n_points + point) * W ..]: raw words, canonical.
template <u32 W>
__global__ void __launch_bounds__(kEvalThreads) poly_eval_kernel(const u64 *__restrict__ polys, u64 n,
                                                                  const u64 *__restrict__ points_val, u32 n_points,
                                                                  u64 *__restrict__ out) {
    __shared__ u64 sh[kEvalThreads * W];
    const u32 tid = threadIdx.x;
    const u32 point = blockIdx.x;
    const u64 poly = blockIdx.y;
    const u64 *coeffs = polys + poly * n * W;
    EvalElem<W> x;
#pragma unroll
    for (u32 k = 0; k < W; k++) x.c[k] = points_val[point * W + k];
    // thread t owns coefficients [t * chunk, (t + 1) * chunk): Horner from the top of its chunk
    const u64 chunk = (n + kEvalThreads - 1) / kEvalThreads;
    const u64 lo = (u64)tid * chunk;
    const u64 hi = lo + chunk < n ? lo + chunk : n;
    EvalElem<W> acc;
#pragma unroll
    for (u32 k = 0; k < W; k++) acc.c[k] = 0;
    for (u64 i = hi; i > lo; i--) {
        EvalElem<W> ci;
#pragma unroll
        for (u32 k = 0; k < W; k++) ci.c[k] = coeffs[(i - 1) * W + k];
        acc = eval_add<W>(eval_mul<W>(acc, x), ci);
    }
    // y = x^chunk (values, so products of values stay values)
    EvalElem<W> y, base = x;
#pragma unroll
    for (u32 k = 0; k < W; k++) y.c[k] = k == 0 ? 1ull : 0ull;
    for (u64 e = chunk; e; e >>= 1) {
        if (e & 1) y = eval_mul<W>(y, base);
        base = eval_mul<W>(base, base);
    }
    // tree: P_t <- P_t + P_(t + s) * y^s, s = 1, 2, 4, ..
    for (u32 s = 1; s < kEvalThreads; s <<= 1) {
#pragma unroll
        for (u32 k = 0; k < W; k++) sh[tid * W + k] = acc.c[k];
        __syncthreads();
        if ((tid & (2 * s - 1)) == 0) {
            EvalElem<W> other;
#pragma unroll
            for (u32 k = 0; k < W; k++) other.c[k] = sh[(tid + s) * W + k];
            acc = eval_add<W>(acc, eval_mul<W>(other, y));
        }
        y = eval_mul<W>(y, y);
        __syncthreads();
    }
    if (tid == 0) {
#pragma unroll
        for (u32 k = 0; k < W; k++) out[(poly * n_points + point) * W + k] = acc.c[k];
    }
}

// ---- window update of Polynomial::reduce_by_ntt_friendly_modulus (polynomial.rs:1126-1143) ---------------
// new[i] = (i < chunk ? coeffs[chunk_index * chunk + i] : old[i - chunk]) - product[i], element-wise on
// w-word elements; `old` holds the low tail_length elements of the previous window.  Canonical output.
__global__ void reduce_window_kernel(const u64 *__restrict__ coeffs_chunk, const u64 *__restrict__ old_tail,
                                     const u64 *__restrict__ product, u64 chunk_words, u64 total_words,
                                     u64 *__restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_words) return;
    const u64 x = i < chunk_words ? coeffs_chunk[i] : old_tail[i - chunk_words];
    out[i] = gl_sub(gl_canon(x), gl_canon(product[i]));
}

// ---- point-wise division for Polynomial::clean_divide (polynomial.rs:2358-2413) ---------------------------
// q[i] = a[i] / b[i] on raw Montgomery words: raw(a / b) = a_raw * (b_raw)^-1 * 2^64 (plain mod-p arithmetic),
// the inverse by Fermat (b^(p-2), p - 2 = 2^64 - 2^32 - 1).  A zero divisor value raises *flag (the coset hit a
// root of the divisor; the host retries with another offset).  The result is lazy (any u64).
__global__ void pointwise_divide_kernel(u64 *__restrict__ a, const u64 *__restrict__ b, u64 n, u32 *flag) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 x = gl_canon(b[i]);
    if (x == 0) {
        *flag = 1;
        return;
    }
    // p - 2 = 0xFFFFFFFE_FFFFFFFF = (2^32 - 2) * 2^32 + (2^32 - 1):
    //   f = x^(2^32 - 2) = (x^(2^31 - 1))^2,  x^(p-2) = f^(2^32) * (f * x)
    u64 acc = x;  // x^(2^1 - 1)
#pragma unroll 1
    for (int k = 0; k < 30; k++) acc = gl_mul(gl_mul(acc, acc), x);  // x^(2^31 - 1)
    const u64 f = gl_mul(acc, acc);
    u64 hi = f;
#pragma unroll 1
    for (int k = 0; k < 32; k++) hi = gl_mul(hi, hi);
    const u64 inv = gl_mul(hi, gl_mul(f, x));
    a[i] = gl_mul(gl_mul(a[i], inv), GL_EPS);
}

// ---- host planning ------------------------------------------------------------------------------

inline size_t col_pass_smem(u32 log_np) {
    u32 np = 1u << log_np;
    return ((size_t)(np + (np >> 5) + 1) * kNttColStride + (np >> 1)) * sizeof(u64);
}
inline size_t row_pass_smem(u32 log_nk, u32 to, u32 w) {
    u32 nk = 1u << log_nk;
    u32 stride = (to * w) | 1;
    return ((size_t)(nk + (nk >> 5) + 1) * stride + (nk >> 1)) * sizeof(u64);
}
inline u32 pick_threads(u64 items) {
    u32 nt = 64;
    while (nt < 1024 && (u64)nt * 8 < items) nt <<= 1;
    return nt;
}

struct NttPlan {
    u32 k;
    u32 l[4];
};

inline NttPlan make_plan(u32 log_n) {
    NttPlan p{};
    p.k = (log_n + kNttMaxLogPass - 1) / kNttMaxLogPass;
    if (p.k == 0) p.k = 1;
    for (u32 i = 0; i < p.k; i++) p.l[i] = log_n / p.k + (i < log_n % p.k ? 1 : 0);
    return p;
}

}  // namespace tf21
