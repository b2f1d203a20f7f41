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
//     (gl_shlc / gl_canonw), sums and differences stay "any u64" (gl_addl / gl_sub), so a butterfly
//     is 8 ALU instructions + 1 IMAD.WIDE instead of 12 ALU instructions;
//   * one general multiply per element between the two steps (omega_1024^(b k1)) and, for column
//     passes, one for the inter-pass twiddle omega_B^(i j_rest), read from a full [j_rest][i] table in
//     L2 (coalesced, 8 MiB for B = 2^20).  Both steps run through the same loop body (halves the
//     instruction footprint: the unrolled 32-point DFT is ~1400 instructions);
//   * the 32 x 32 transpose between the steps goes through the warp's own slice of shared memory
//     (__syncwarp only); __syncthreads is needed only around the coalesced global staging.
//
// Values in flight are lazy (any u64 representative); only the last pass canonicalises on store, so
// non-canonical input words are accepted as well.
//
// Shared-memory layout: tile[col][S] u64 with S = 1058 (= 2 mod 16: conflict-free for both the
// 8-lanes-per-row staging pattern and the per-warp column patterns).
#pragma once
#include "ntt_kernels.cuh"

namespace tf21 {

constexpr u32 kFastCols = 8;        // word-columns (= warps) per CTA
constexpr u32 kFastS = 1058;        // u64 per column slice in shared memory (>= 32 * 33)
constexpr u32 kFastThreads = kFastCols * 32;
constexpr size_t kFastSmem = (size_t)kFastCols * kFastS * sizeof(u64);

__host__ __device__ constexpr u32 brev5(u32 k) {
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// 32-point DFT in registers, decimation in time: in v[a] = x[a] (any u64), out v[brev5(k)] =
// sum_a x[a] w^(a k) (any u64), w = omega_32^(+-1) = 2^(+-78).  (butterfly of ntt.rs:203-210;
// the twiddles are shifts; exponents >= 96 use 2^96 = -1 and swap the roles of sum and difference)
template <bool INV>
__device__ __forceinline__ void dft32(u64 (&v)[32]) {
#pragma unroll
    for (int ls = 1; ls <= 5; ls++) {
        const int m = 1 << ls, half = m >> 1;
#pragma unroll
        for (int k = 0; k < 32; k += m) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const int iu = brev5(k + j), ib = brev5(k + j + half);
                int E = (78 * j * (32 / m)) % 192;
                if (INV) E = (192 - E) % 192;
                const bool neg = E >= 96;
                const int S = neg ? E - 96 : E;
                const u64 t = (S == 0) ? gl_canonw(v[ib]) : gl_shlc(v[ib], S);
                const u64 u = v[iu];
                if (!neg) {
                    v[iu] = gl_addl(u, t);
                    v[ib] = gl_sub(u, t);
                } else {
                    v[iu] = gl_sub(u, t);
                    v[ib] = gl_addl(u, t);
                }
            }
        }
    }
}

// The 1024-point DFT of the column held by this warp.
// in : v[a] = x[32 a + lane]                       (natural order, any u64)
// out: slice[i] = X[i] * (tw1 ? tw1[32 (i >> 5)] : 1), i = lane + 32 k2   (any u64)
// slice: this warp's kFastS-word shared-memory slice; tw0 = t1 + lane with t1[k1*32+b] = w1024^(k1 b);
// tw1: per-lane pointer into the inter-pass twiddle row (or nullptr).
template <bool INV>
__device__ __forceinline__ void dft1024_warp(u64 (&v)[32], u64 *slice, const u64 *tw0, const u64 *tw1, u32 lane) {
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        dft32<INV>(v);
        const u64 *tw = it ? tw1 : tw0;
        u64 *out = slice + lane;
        const u32 ss = it ? 32u : 33u;
        __syncwarp();
        if (tw) {
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = gl_mul(v[brev5(k)], __ldg(tw + 32 * k));
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) out[k * ss] = v[brev5(k)];
        }
        __syncwarp();
        if (it == 0) {
#pragma unroll
            for (int b = 0; b < 32; b++) v[b] = slice[lane * 33 + b];
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

template <bool INV>
__global__ void __launch_bounds__(kFastThreads, 2) ntt1024_col_kernel(const FastColArgs a) {
    extern __shared__ u64 smem[];
    u64 *tile = smem;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 ct = blockIdx.x % a.n_col_tiles;
    const u32 rest = blockIdx.x / a.n_col_tiles;
    const u32 o = rest % a.n_outer, b = rest / a.n_outer;
    const u64 q0 = (u64)ct * kFastCols;
    const u64 inner_elems = a.inner_words / a.w;
    const u32 c = lane & 7, rsub = lane >> 3;

    // ---- stage in: 8 lanes per 64-byte row segment, 4 rows per warp instruction ----
    const u64 block_off = (u64)o * 1024 * a.inner_words + q0;
    {
        const u64 *src = a.src + (u64)b * a.src_array_words + block_off + c;
        const u64 jcol = (q0 + c) / a.w;
        u64 *tl = tile + c * kFastS;
#pragma unroll 8
        for (u32 it = 0; it < 32; it++) {
            const u32 r = warp * 128 + it * 4 + rsub;
            const u64 j = ((u64)o * 1024 + r) * inner_elems + jcol;
            u64 x = 0;
            if (j < a.n_in_elems) {
                x = src[(u64)r * a.inner_words];
                if (a.pre.lo) x = gl_mul(x, scale_factor(a.pre, j));
            }
            tl[r] = x;
        }
    }
    __syncthreads();

    // ---- the column of this warp ----
    u64 *slice = tile + warp * kFastS;
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) v[aa] = slice[32 * aa + lane];
    const u64 jrest = (q0 + warp) / a.w;
    // inter-pass twiddle omega_B^(i * j_rest), i = lane + 32 k2: straight from the full table when there is one
    dft1024_warp<INV>(v, slice, a.t1 + lane, a.tw_full ? a.tw_full + jrest * 1024 + lane : nullptr, lane);
    if (!a.tw_full) {
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
    {
        u64 *dst = a.dst + (u64)b * a.dst_array_words + block_off + c;
        const u64 *tl = tile + c * kFastS;
#pragma unroll 8
        for (u32 it = 0; it < 32; it++) {
            const u32 r = warp * 128 + it * 4 + rsub;
            dst[(u64)r * a.inner_words] = tl[r];
        }
    }
}

struct FastRowArgs {
    const u64 *src;
    u64 *dst;
    u64 array_words;      // n (w = 1)
    u32 n_tiles_t, mid;   // tiles of 8 consecutive i_1; number of mid values
    u64 src_t_stride;     // words between consecutive i_1 rows
    u64 dst_i_stride, dst_mid_stride;
    const u64 *t1;
    u64 post_scalar;      // 0 => none
    ScaleTab post;
    u64 elem_i_stride, elem_mid_stride;
};

// Last pass of a multi-pass transform, w = 1: rows are contiguous (direct coalesced loads), the
// output is transposed (consecutive i_1 adjacent) and goes through shared memory; canonical on store.
template <bool INV>
__global__ void __launch_bounds__(kFastThreads, 2) ntt1024_row_kernel(const FastRowArgs a) {
    extern __shared__ u64 smem[];
    u64 *tile = smem;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 tt = blockIdx.x % a.n_tiles_t;
    const u32 rest = blockIdx.x / a.n_tiles_t;
    const u32 mid = rest % a.mid, b = rest / a.mid;
    const u32 t0 = tt * kFastCols;

    const u64 *row = a.src + (u64)b * a.array_words + (u64)(t0 + warp) * a.src_t_stride + (u64)mid * 1024;
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) v[aa] = row[32 * aa + lane];
    u64 *slice = tile + warp * kFastS;
    dft1024_warp<INV>(v, slice, a.t1 + lane, nullptr, lane);
    __syncthreads();

    // stage out: element (i_1 = t0 + c, i_k = r) -> dst[r * dst_i_stride + (t0 + c)]
    {
        const u32 c = lane & 7, rsub = lane >> 3;
        u64 *dst = a.dst + (u64)b * a.array_words + (u64)mid * a.dst_mid_stride + t0 + c;
        const u64 *tl = tile + c * kFastS;
        const u64 elem_base = (u64)(t0 + c) + (u64)mid * a.elem_mid_stride;
#pragma unroll 8
        for (u32 it = 0; it < 32; it++) {
            const u32 r = warp * 128 + it * 4 + rsub;
            u64 x = tl[r];
            if (a.post_scalar) x = gl_mul(x, a.post_scalar);
            if (a.post.lo) x = gl_mul(x, scale_factor(a.post, elem_base + (u64)r * a.elem_i_stride));
            dst[(u64)r * a.dst_i_stride] = gl_canonw(x);
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

// n = 1024, w = 1: one warp per array of the batch, no global staging at all.
template <bool INV>
__global__ void __launch_bounds__(kFastThreads, 2) ntt1024_single_kernel(const FastSingleArgs a) {
    extern __shared__ u64 smem[];
    u64 *tile = smem;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 arr = (u64)blockIdx.x * kFastCols + warp;
    if (arr >= a.batch) return;
    const u64 *row = a.src + arr * a.src_array_words;
    u64 v[32];
#pragma unroll
    for (int aa = 0; aa < 32; aa++) {
        const u64 j = 32 * aa + lane;
        u64 x = 0;
        if (j < a.n_in_elems) {
            x = row[j];
            if (a.pre.lo) x = gl_mul(x, scale_factor(a.pre, j));
        }
        v[aa] = x;
    }
    u64 *slice = tile + warp * kFastS;
    dft1024_warp<INV>(v, slice, a.t1 + lane, nullptr, lane);
    u64 *out = a.dst + arr * a.dst_array_words;
#pragma unroll 8
    for (int k2 = 0; k2 < 32; k2++) {
        u64 x = slice[lane + 32 * k2];
        if (a.post_scalar) x = gl_mul(x, a.post_scalar);
        if (a.post.lo) x = gl_mul(x, scale_factor(a.post, (u64)(lane + 32 * k2)));
        out[lane + 32 * k2] = gl_canonw(x);
    }
}


// ---- tables for the fast path ---------------------------------------------------------------------
struct FastTables {
    std::map<int, u64 *> t1;                                   // inverse -> [32][32]
    std::map<std::tuple<unsigned, int, u64>, u64 *> tw_full;   // (log_b, inverse, scalar) -> [B/1024][1024]
};
static std::map<int, FastTables> g_fast_tables;  // by device, guarded by g_mutex

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

// T[j][i] = scalar * omega_B^(+-j i), j < B/1024, i < 1024
inline int get_tw_full(DeviceTables &t, int dev, unsigned log_b, int inverse, u64 scalar, const u64 **out) {
    FastTables &ft = g_fast_tables[dev];
    auto key = std::make_tuple(log_b, inverse, scalar);
    auto it = ft.tw_full.find(key);
    if (it != ft.tw_full.end()) {
        *out = it->second;
        return 0;
    }
    u64 w = hgl_root_of_unity(log_b);
    if (inverse) w = hgl_inv(w);
    const u64 rows = (1ull << log_b) >> 10;
    std::vector<u64> h(rows * 1024);
    u64 step = 1;  // w^j
    for (u64 j = 0; j < rows; j++) {
        u64 acc = scalar % GL_P;
        for (u32 i = 0; i < 1024; i++) {
            h[j * 1024 + i] = acc;
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

template <typename K, typename A>
inline int launch_fast(K kernel, unsigned grid, const A &args, cudaStream_t st) {
    TF21_LAUNCH(kernel, grid, kFastThreads, kFastSmem, st, args);
    return 0;
}

// Core entry: dst[b] = scale_post( NTT_n( zero_extend( scale_pre( src[b][0..n_in) ) ) ) ) for b < batch.
// src arrays are n_in*w words apart, dst arrays n*w words apart.  src == dst is allowed when
// n_in == n.  `scratch` (n*w*batch words) is required when log2 n > 10.  Caller holds no lock;
// table lookups lock g_mutex internally.
inline int ntt_run(DeviceTables &tabs, int dev, const u64 *src, u64 n_in, u64 *dst, u64 n, u32 w, u64 batch,
                   int inverse, ScaleTab pre, ScaleTab post, u64 post_scalar, u64 *scratch, cudaStream_t st) {
    const u32 log_n = ilog2_u64(n);
    const NttPlan plan = make_plan(log_n);
    const u64 array_words = n * w;
    const u64 *tw_small = tabs.tw_small[inverse ? 1 : 0];
    const u64 *t1 = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        TF21_TRY(get_t1(tabs, dev, inverse, &t1));
    }

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
        if (lp == 10) {
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
            a.log_b = log_b;
            a.pre = cur_pre;
            {
                std::lock_guard<std::mutex> lock(g_mutex);
                if (log_b <= kFullTwiddleMaxLog) {
                    u64 scalar = 1;
                    if (p == 0 && post_scalar) {  // fold the unscale (ntt.rs:220-228) into pass 1
                        scalar = post_scalar;
                        post_scalar = 0;
                    }
                    TF21_TRY(get_tw_full(tabs, dev, log_b, inverse, scalar, &a.tw_full));
                } else {
                    DeviceTables::Split sp;
                    TF21_TRY(get_split_tables(tabs, log_b, inverse, &sp));
                    a.tw = ScaleTab{sp.lo, sp.hi, sp.h};
                }
            }
            u64 grid = batch * n_outer * a.n_col_tiles;
            if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
            if (inverse)
                TF21_TRY(launch_fast(ntt1024_col_kernel<true>, (unsigned)grid, a, st));
            else
                TF21_TRY(launch_fast(ntt1024_col_kernel<false>, (unsigned)grid, a, st));
        } else {
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
        }
        consumed += lp;
        cur_src = scratch;
        cur_src_words = array_words;
        cur_n_in = n;
        cur_pre = ScaleTab{nullptr, nullptr, 0};
    }

    const u32 lk = plan.l[plan.k - 1];
    const u32 nk = 1u << lk;
    if (lk == 10 && w == 1 && plan.k == 1) {
        FastSingleArgs a{};
        a.src = cur_src;
        a.dst = dst;
        a.src_array_words = cur_src_words;
        a.dst_array_words = array_words;
        a.batch = batch;
        a.n_in_elems = cur_n_in;
        a.t1 = t1;
        a.post_scalar = post_scalar;
        a.pre = cur_pre;
        a.post = post;
        u64 grid = (batch + kFastCols - 1) / kFastCols;
        if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
        if (inverse) return launch_fast(ntt1024_single_kernel<true>, (unsigned)grid, a, st);
        return launch_fast(ntt1024_single_kernel<false>, (unsigned)grid, a, st);
    }
    if (lk == 10 && w == 1 && plan.k >= 2 && plan.l[0] >= 3) {
        FastRowArgs a{};
        const u32 n1 = 1u << plan.l[0];
        const u32 midc = (plan.k == 3) ? (1u << plan.l[1]) : 1u;
        const u64 o_total = n >> lk;
        a.src = cur_src;
        a.dst = dst;
        a.array_words = array_words;
        a.n_tiles_t = n1 / kFastCols;
        a.mid = midc;
        a.src_t_stride = (u64)midc * 1024;
        a.dst_i_stride = o_total;
        a.dst_mid_stride = n1;
        a.t1 = t1;
        a.post_scalar = post_scalar;
        a.post = post;
        a.elem_i_stride = o_total;
        a.elem_mid_stride = n1;
        u64 grid = batch * a.n_tiles_t * a.mid;
        if (grid > 0x7fffffffull) return TF21_E_LEN_TOO_LARGE;
        if (inverse) return launch_fast(ntt1024_row_kernel<true>, (unsigned)grid, a, st);
        return launch_fast(ntt1024_row_kernel<false>, (unsigned)grid, a, st);
    }

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

}  // namespace tf21
