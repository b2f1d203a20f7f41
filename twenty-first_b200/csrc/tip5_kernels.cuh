// tip5_kernels.cuh -- batched Tip5 kernels and the level-batched Merkle build (K3, K4 of SURVEY.md).
//
// Replaces Tip5::{permutation, hash_10, hash_pair, hash_varlen} (tip5/mod.rs:529-533, 559-586,
// 617-623; sponge.rs:41-56) applied over a batch -- the caller-side `par_iter().map(Tip5::hash_10)`
// pattern of benches/tip5.rs:43-49 -- and MerkleTree::{par_new, sequential_new, *_frugal_root}
// (util_types/merkle_tree.rs:149-222, 299-364, 393-429).
//
// Data layout in HBM: digests are 5 consecutive u64 (40 B), a hash_10 input is 10 consecutive u64
// (80 B, 16-byte aligned), a Merkle node array is heap indexed exactly like the reference's
// `Vec<Digest>`: nodes[0] = 0, nodes[1] = root, nodes[n..2n) = leaves.  The children of the level
// holding `cnt` nodes are the 10*cnt contiguous words at nodes + 10*cnt, so one level is one
// hash_10 batch whose output is written at nodes + 5*cnt.
#pragma once
#include "runtime.cuh"
#include "tip5.cuh"

namespace tf21 {

// ROUND_CONSTANTS, tip5/mod.rs:68-149 (canonical values; the kernel uses value * 2^64 mod p)
static const u64 kTip5RoundConstants[TIP5_ROUNDS * TIP5_STATE] = {
    13630775303355457758ull, 16896927574093233874ull, 10379449653650130495ull, 1965408364413093495ull,
    15232538947090185111ull, 15892634398091747074ull, 3989134140024871768ull,  2851411912127730865ull,
    8709136439293758776ull,  3694858669662939734ull,  12692440244315327141ull, 10722316166358076749ull,
    12745429320441639448ull, 17932424223723990421ull, 7558102534867937463ull,  15551047435855531404ull,
    17532528648579384106ull, 5216785850422679555ull,  15418071332095031847ull, 11921929762955146258ull,
    9738718993677019874ull,  3464580399432997147ull,  13408434769117164050ull, 264428218649616431ull,
    4436247869008081381ull,  4063129435850804221ull,  2865073155741120117ull,  5749834437609765994ull,
    6804196764189408435ull,  17060469201292988508ull, 9475383556737206708ull,  12876344085611465020ull,
    13835756199368269249ull, 1648753455944344172ull,  9836124473569258483ull,  12867641597107932229ull,
    11254152636692960595ull, 16550832737139861108ull, 11861573970480733262ull, 1256660473588673495ull,
    13879506000676455136ull, 10564103842682358721ull, 16142842524796397521ull, 3287098591948630584ull,
    685911471061284805ull,   5285298776918878023ull,  18310953571768047354ull, 3142266350630002035ull,
    549990724933663297ull,   4901984846118077401ull,  11458643033696775769ull, 8706785264119212710ull,
    12521758138015724072ull, 11877914062416978196ull, 11333318251134523752ull, 3933899631278608623ull,
    16635128972021157924ull, 10291337173108950450ull, 4142107155024199350ull,  16973934533787743537ull,
    11068111539125175221ull, 17546769694830203606ull, 5315217744825068993ull,  4609594252909613081ull,
    3350107164315270407ull,  17715942834299349177ull, 9600609149219873996ull,  12894357635820003949ull,
    4597649658040514631ull,  7735563950920491847ull,  1663379455870887181ull,  13889298103638829706ull,
    7375530351220884434ull,  3502022433285269151ull,  9231805330431056952ull,  9252272755288523725ull,
    10014268662326746219ull, 15565031632950843234ull, 1209725273521819323ull,  6024642864597845108ull,
};

inline int upload_tip5_constants() {
    double lo[TIP5_ROUNDS * TIP5_STATE], hi[TIP5_ROUNDS * TIP5_STATE];
    for (int i = 0; i < TIP5_ROUNDS * TIP5_STATE; i++) {
        u64 raw = hgl_to_raw(kTip5RoundConstants[i]);
        lo[i] = (double)(raw & 0xffffffffull);  // 32-bit halves: exact in a double
        hi[i] = (double)(raw >> 32);
    }
    // fixed-length round 0: lanes 10..15 = ONE (raw 2^32 - 1); S-box there is ONE^7 on the raw word
    double lo0[TIP5_STATE], hi0[TIP5_STATE];
    {
        const u64 c7 = hgl_pow(TIP5_RAW_ONE, 7);
        const u64 c7_lo = c7 & 0xffffffffull, c7_hi = c7 >> 32;
        static const u32 mds[16] = {61402, 1108, 28750, 33823, 7454, 43244, 53865, 12034,
                                    56951, 27521, 41351, 40901, 12021, 59689, 26798, 17845};
        for (int i = 0; i < TIP5_STATE; i++) {
            u64 sl = 0, sh = 0;
            for (int j = TIP5_RATE; j < TIP5_STATE; j++) {
                sl += (u64)mds[(i - j) & 15] * c7_lo;
                sh += (u64)mds[(i - j) & 15] * c7_hi;
            }
            lo0[i] = lo[i] + (double)sl;  // both < 2^51: exact
            hi0[i] = hi[i] + (double)sh;
        }
    }
#if TIP5_MDS_CRT
    // seeds of the cyclic / negacyclic half-size products (tip5.cuh): halves of sums and differences, exact
    auto to_seeds = [](double *v) {
        for (int i = 0; i < 8; i++) {
            const double a = v[i], b = v[i + 8];
            v[i] = 0.5 * (a + b);
            v[i + 8] = 0.5 * (a - b);
        }
    };
    to_seeds(lo0);
    to_seeds(hi0);
    for (int r = 0; r < TIP5_ROUNDS; r++) {
        to_seeds(lo + r * TIP5_STATE);
        to_seeds(hi + r * TIP5_STATE);
    }
#endif
    {
        u64 raw[TIP5_ROUNDS * TIP5_STATE];
        for (int i = 0; i < TIP5_ROUNDS * TIP5_STATE; i++) raw[i] = hgl_to_raw(kTip5RoundConstants[i]);
        TF21_CUDA(cudaMemcpyToSymbol(c_tip5_rc_raw, raw, sizeof(raw)));
    }
    TF21_CUDA(cudaMemcpyToSymbol(c_tip5_rc0f_lo, lo0, sizeof(lo0)));
    TF21_CUDA(cudaMemcpyToSymbol(c_tip5_rc0f_hi, hi0, sizeof(hi0)));
    uint8_t lut[256];
    // LOOKUP_TABLE, tip5/mod.rs:50-64 = ((x+1)^3 + 256) mod 257 (tip5/mod.rs:1022-1053)
    for (unsigned i = 0; i < 256; i++) lut[i] = (uint8_t)(((i + 1) * (i + 1) * (i + 1) + 256) % 257);
    TF21_CUDA(cudaMemcpyToSymbol(c_tip5_rc_lo, lo, sizeof(lo)));
    TF21_CUDA(cudaMemcpyToSymbol(c_tip5_rc_hi, hi, sizeof(hi)));
    TF21_CUDA(cudaMemcpyToSymbol(c_tip5_lut, lut, sizeof(lut)));
    TF21_CUDA(cudaStreamSynchronize(nullptr));  // pageable sources: drain before kernels on non-blocking streams read them
    return 0;
}

#ifndef TF21_MERKLE_COOP_CNT
#define TF21_MERKLE_COOP_CNT 4096
#endif
constexpr u32 kMerkleCoopCnt = TF21_MERKLE_COOP_CNT;  // batches (Merkle levels, rows) with <= this many nodes use 16 lanes per hash (latency bound)
constexpr int kTip5Threads = 128;
#ifndef TIP5_MIN_BLOCKS
#define TIP5_MIN_BLOCKS 4  /* resident CTAs per SM the register allocation aims for (A/B: tools/ab.sh) */
#endif

// in-place permutation of `count` 16-word states (Tip5::permutation, tip5/mod.rs:529-533)
__global__ void __launch_bounds__(kTip5Threads) tip5_permute_kernel(u64 *__restrict__ states, u64 count) {
    __shared__ uint8_t s_lut[256];
    tip5_load_lut(s_lut);
    __syncthreads();
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[TIP5_STATE];
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(states + 16 * i);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        ulonglong2 v = src[k];
        s[2 * k] = v.x;
        s[2 * k + 1] = v.y;
    }
    tip5_permutation(s, s_lut);
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(states + 16 * i);
#pragma unroll
    for (int k = 0; k < 8; k++) dst[k] = make_ulonglong2(gl_canon(s[2 * k]), gl_canon(s[2 * k + 1]));
}

#ifndef TIP5_HASH10_FIXED
#define TIP5_HASH10_FIXED 1
#endif
// Tip5::hash_10 / hash_pair over a batch (tip5/mod.rs:559-586): state = in[0..10) | ONE x 6.
// COPY: also store the 10 input words at copy_dst + 10 i -- the leaf level of the Merkle build, where the
// reference copies the leaves into the node array (merkle_tree.rs:426) before hashing them.
template <bool COPY>
__global__ void __launch_bounds__(kTip5Threads, TIP5_MIN_BLOCKS)
    tip5_hash10_kernel(const u64 *__restrict__ in, u64 count, u64 *__restrict__ out, u64 *__restrict__ copy_dst) {
    __shared__ uint8_t s_lut[256];
    tip5_load_lut(s_lut);
    __syncthreads();
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[TIP5_STATE];
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + 10 * i);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        ulonglong2 v = src[k];
        s[2 * k] = v.x;
        s[2 * k + 1] = v.y;
        if (COPY) reinterpret_cast<ulonglong2 *>(copy_dst + 10 * i)[k] = v;
    }
#if TIP5_HASH10_FIXED
    tip5_permutation<true, true>(s, s_lut);  // capacity lanes = ONE, folded into the round-0 constants; digest lanes only
#else
#pragma unroll
    for (int k = TIP5_RATE; k < TIP5_STATE; k++) s[k] = TIP5_RAW_ONE;
    tip5_permutation<false, true>(s, s_lut);  // one copy of the round in the instruction stream
#endif
    u64 *dst = out + 5 * i;
#pragma unroll
    for (int k = 0; k < TIP5_DIGEST; k++) dst[k] = gl_canon(s[k]);
}

// ---- cooperative kernels for the top of a Merkle tree (tip5_permutation_coop, tip5.cuh) ----------------
__device__ __forceinline__ void tip5_coop_setup(uint8_t *s_lut, u64 *s_rc) {
    tip5_load_lut(s_lut);
    for (int i = threadIdx.x; i < TIP5_ROUNDS * TIP5_STATE; i += blockDim.x) s_rc[i] = c_tip5_rc_raw[i];
}

// hash_pair of node pair i of a level: 16 lanes, lane l < 10 loads word l, lanes 10..15 hold ONE
__device__ __forceinline__ void tip5_coop_hash10(const u64 *__restrict__ in, u64 *__restrict__ out, u32 lane16,
                                                 const uint8_t *s_lut, const u64 *s_rc) {
    u64 s = lane16 < TIP5_RATE ? in[lane16] : TIP5_RAW_ONE;
    s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
    if (lane16 < TIP5_DIGEST) out[lane16] = s;
}

constexpr int kCoopThreads = 128;  // 8 hashes per CTA
__global__ void __launch_bounds__(kCoopThreads) tip5_hash10_coop_kernel(const u64 *__restrict__ in, u64 count,
                                                                        u64 *__restrict__ out) {
    __shared__ uint8_t s_lut[256];
    __shared__ u64 s_rc[TIP5_ROUNDS * TIP5_STATE];
    tip5_coop_setup(s_lut, s_rc);
    __syncthreads();
    const u64 h = ((u64)blockIdx.x * kCoopThreads + threadIdx.x) >> 4;
    // whole 16-lane groups leave together (count is not necessarily a multiple of 8), the shuffles inside a
    // group only name lanes of the group
    if (h >= count) return;
    tip5_coop_hash10(in + 10 * h, out + 5 * h, threadIdx.x & 15, s_lut, s_rc);
}

// Top of a Merkle tree in one CTA: levels with cnt = first_cnt, first_cnt/2, ..., 1.
// nodes: heap-indexed array; the children level (2*first_cnt nodes) is already complete.
constexpr int kMerkleTailThreads = 1024;  // 64 cooperative hashes in flight
__global__ void __launch_bounds__(kMerkleTailThreads) merkle_tail_kernel(u64 *nodes, u32 first_cnt) {
    __shared__ uint8_t s_lut[256];
    __shared__ u64 s_rc[TIP5_ROUNDS * TIP5_STATE];
    tip5_coop_setup(s_lut, s_rc);
    __syncthreads();
    const u32 group = threadIdx.x >> 4, lane16 = threadIdx.x & 15;
    for (u32 cnt = first_cnt; cnt >= 1; cnt >>= 1) {
        for (u32 i = group; i < cnt; i += kMerkleTailThreads / 16)
            tip5_coop_hash10(nodes + 10ull * (cnt + i), nodes + 5ull * (cnt + i), lane16, s_lut, s_rc);
        __threadfence_block();
        __syncthreads();
    }
}

// Top of a Merkle tree in ONE persistent kernel: every level with cnt = first_cnt, first_cnt/2, ..., 1 nodes, the
// levels separated by a grid-wide barrier instead of a kernel boundary.  One CTA per SM (cooperative launch: all CTAs
// are resident, so the spin barrier cannot deadlock), 16 lanes per hash, consecutive hashes on different SMs so that a
// level with few nodes runs one 16-lane group per SM at the lowest latency the permutation has.  Replaces ~16 launches
// (two thread-per-hash levels, five cooperative levels, the single-CTA tail) of the level-batched build
// (sequentially_fill_tree, merkle_tree.rs:216-222): each level costs the latency of one cooperative hash plus a barrier
// instead of a launch, which is what a 2^16-leaf tree (a FRI round) consists of.
// Children digests were written by other SMs earlier in this kernel: they are read with ld.global.cg (L2), never
// through the non-coherent path.  bar: a zeroed u32, counts arrivals (monotonic over the levels).
constexpr int kMerkleTopThreads = 1024;
__global__ void __launch_bounds__(kMerkleTopThreads, 1) merkle_top_kernel(u64 *nodes, u32 first_cnt, u32 *bar) {
    __shared__ uint8_t s_lut[256];
    __shared__ u64 s_rc[TIP5_ROUNDS * TIP5_STATE];
    tip5_coop_setup(s_lut, s_rc);
    __syncthreads();
    const u32 lane16 = threadIdx.x & 15;
    const u32 g = (threadIdx.x >> 4) * gridDim.x + blockIdx.x;
    const u32 n_groups = gridDim.x * (kMerkleTopThreads / 16);
    u32 target = 0;
    for (u32 cnt = first_cnt; cnt >= 1; cnt >>= 1) {
        for (u32 i = g; i < cnt; i += n_groups) {
            const u64 *in = nodes + 10ull * (cnt + i);
            u64 s = lane16 < TIP5_RATE ? __ldcg(in + lane16) : TIP5_RAW_ONE;
            s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
            if (lane16 < TIP5_DIGEST) __stcg(nodes + 5ull * (cnt + i) + lane16, s);
        }
        if (cnt == 1) break;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();  // this CTA's digests are visible device-wide before its arrival is
            atomicAdd(bar, 1u);
            target += gridDim.x;
            while (*(volatile u32 *)bar < target) {
            }
            __threadfence();
        }
        __syncthreads();
    }
}

// Tip5::hash_varlen over rows (tip5/mod.rs:617-623, sponge.rs:41-56): one row per thread,
// overwrite-mode absorb of 10-word chunks, padding 1,0,.. always present.  Element k of row i is
// data[i * row_stride + k * elem_stride]: (row_len, 1) for a row-major matrix, (1, column stride)
// for column-major data such as a batch of NTT codewords (coalesced across the threads of a warp).
__global__ void __launch_bounds__(kTip5Threads)
    tip5_hash_rows_kernel(const u64 *__restrict__ data, u64 row_len, u64 n_rows, u64 row_stride, u64 elem_stride,
                          u64 *__restrict__ out) {
    __shared__ uint8_t s_lut[256];
    tip5_load_lut(s_lut);
    __syncthreads();
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const u64 *row = data + i * row_stride;
    u64 s[TIP5_STATE];
#pragma unroll
    for (int k = 0; k < TIP5_STATE; k++) s[k] = 0;
    u64 full = row_len / TIP5_RATE;
    for (u64 c = 0; c < full; c++) {
#pragma unroll
        for (int k = 0; k < TIP5_RATE; k++) s[k] = row[(c * TIP5_RATE + k) * elem_stride];
        tip5_permutation(s, s_lut);
    }
    u32 rem = (u32)(row_len - full * TIP5_RATE);
#pragma unroll
    for (int k = 0; k < TIP5_RATE; k++) {
        u64 v = 0;
        if ((u32)k < rem) v = row[(full * TIP5_RATE + k) * elem_stride];
        if ((u32)k == rem) v = TIP5_RAW_ONE;
        s[k] = v;
    }
    tip5_permutation(s, s_lut);
    u64 *dst = out + 5 * i;
#pragma unroll
    for (int k = 0; k < TIP5_DIGEST; k++) dst[k] = gl_canon(s[k]);
}

// Cooperative form of the same sponge for few long rows (Tip5::hash_varlen of one sequence is a chain of
// dependent permutations): 16 lanes per row, lane l < 10 overwrites its rate element per chunk.
__global__ void __launch_bounds__(kCoopThreads)
    tip5_hash_rows_coop_kernel(const u64 *__restrict__ data, u64 row_len, u64 n_rows, u64 row_stride, u64 elem_stride,
                               u64 *__restrict__ out) {
    __shared__ uint8_t s_lut[256];
    __shared__ u64 s_rc[TIP5_ROUNDS * TIP5_STATE];
    tip5_coop_setup(s_lut, s_rc);
    __syncthreads();
    const u64 i = ((u64)blockIdx.x * kCoopThreads + threadIdx.x) >> 4;
    if (i >= n_rows) return;
    const u32 lane16 = threadIdx.x & 15;
    const u64 *row = data + i * row_stride;
    u64 s = 0;
    const u64 full = row_len / TIP5_RATE;
    for (u64 c = 0; c < full; c++) {
        if (lane16 < TIP5_RATE) s = row[(c * TIP5_RATE + lane16) * elem_stride];
        s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
    }
    const u32 rem = (u32)(row_len - full * TIP5_RATE);
    if (lane16 < TIP5_RATE) {
        u64 v = 0;
        if (lane16 < rem) v = row[(full * TIP5_RATE + lane16) * elem_stride];
        if (lane16 == rem) v = TIP5_RAW_ONE;
        s = v;
    }
    s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
    if (lane16 < TIP5_DIGEST) out[5 * i + lane16] = s;
}

// nodes[0] = 0 and nodes[n..2n) = leafs  (initialize_merkle_tree_nodes, merkle_tree.rs:393-429)
__global__ void merkle_init_kernel(const u64 *__restrict__ leafs, u64 n_leaf_words, u64 *__restrict__ nodes) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    if (i < 5) nodes[i] = 0;
    for (; i < n_leaf_words; i += stride) nodes[n_leaf_words + i] = leafs[i];
}

// local subtree -> global heap positions (see tf21_merkle_scatter_subtree_dev)
__global__ void merkle_scatter_kernel(const u64 *__restrict__ local, u64 n_local, u64 shard, u64 n_shards,
                                      u64 *__restrict__ global) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;  // word index into local nodes, from node 1
    u64 total = (2 * n_local - 1) * 5;
    if (t >= total) return;
    u64 node = 1 + t / 5, lane = t % 5;
    unsigned level = 63u - (unsigned)__clzll((long long)node);  // node in [2^level, 2^(level+1))
    u64 width = 1ull << level;
    u64 g = n_shards * width + shard * width + (node - width);
    global[5 * g + lane] = local[5 * node + lane];
}

// out[k] = nodes[idx[k]]: the gather behind MerkleTree::authentication_structure (merkle_tree.rs:614-622);
// one thread per word of the result
__global__ void merkle_gather_kernel(const u64 *__restrict__ nodes, const u64 *__restrict__ idx, u64 count,
                                     u64 *__restrict__ out) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 5 * count) return;
    out[t] = nodes[5 * idx[t / 5] + t % 5];
}

// bag_peaks (mmr/mmr_accumulator.rs:379-391): a sequential chain of at most 65 hashes -- one thread.
// acc = hash_10(lo32(leaf_count), hi32(leaf_count), 0 x 8) (raw words passed in by the host), then from the
// last peak to the first: acc = hash_pair(peak, acc).
__global__ void __launch_bounds__(32) mmr_bag_peaks_kernel(const u64 *__restrict__ peaks, u32 n_peaks, u64 lc_lo_raw,
                                                           u64 lc_hi_raw, u64 *__restrict__ out) {
    __shared__ uint8_t s_lut[256];
    __shared__ u64 s_rc[TIP5_ROUNDS * TIP5_STATE];
    tip5_coop_setup(s_lut, s_rc);
    __syncthreads();
    if (threadIdx.x >= 16) return;
    const u32 lane16 = threadIdx.x;
    // 16 lanes = the 16 state elements (tip5_permutation_coop): hash_10(lo, hi, 0 x 8), capacity = ONE
    u64 s = lane16 == 0 ? lc_lo_raw : lane16 == 1 ? lc_hi_raw : lane16 < TIP5_RATE ? 0ull : TIP5_RAW_ONE;
    s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
#pragma unroll 1
    for (u32 k = n_peaks; k-- > 0;) {
        // hash_pair(peak, acc): lanes 0..4 <- peak, lanes 5..9 <- acc (= lanes 0..4 of the previous state)
        const u32 lo = __shfl_sync(0xffffu, (u32)s, (lane16 + 11) & 15, 16);
        const u32 hi = __shfl_sync(0xffffu, (u32)(s >> 32), (lane16 + 11) & 15, 16);
        s = lane16 < TIP5_DIGEST ? peaks[5 * k + lane16] : lane16 < TIP5_RATE ? gl_pack(lo, hi) : TIP5_RAW_ONE;
        s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
    }
    if (lane16 < TIP5_DIGEST) out[lane16] = s;
}

// Tip5::sample_indices (tip5/mod.rs:636-656): squeeze (emit state[0..10), then permute) until num_indices
// elements other than BFieldElement::MAX have been seen; index = value as u32 % upper_bound (a power of two).
// A sequential sponge walk -- one thread.  state: 16 raw words, updated in place like `&mut self`.
__global__ void __launch_bounds__(32) tip5_sample_indices_kernel(u64 *__restrict__ state, u32 upper_bound, u64 num_indices,
                                                                 u64 rinv, u32 *__restrict__ out) {
    __shared__ uint8_t s_lut[256];
    __shared__ u64 s_rc[TIP5_ROUNDS * TIP5_STATE];
    tip5_coop_setup(s_lut, s_rc);
    __syncthreads();
    if (threadIdx.x >= 16) return;
    const u32 lane16 = threadIdx.x;
    u64 s = state[lane16];
    u64 produced = 0;
    while (produced < num_indices) {
        // squeeze: lanes 0..9 hold the produced elements, then the state is permuted (tip5/mod.rs:693-698)
        const u64 value = gl_mulc(gl_canon(s), rinv);  // BFieldElement::value(): raw * 2^-64 mod p
        const bool keep = lane16 < TIP5_RATE && value != GL_P - 1;
        const u32 kept = __ballot_sync(0xffffu, keep) & 0x3ffu;
        const u64 pos = produced + __popc(kept & ((1u << lane16) - 1));
        if (keep && pos < num_indices) out[pos] = (u32)value & (upper_bound - 1);
        produced += __popc(kept);
        s = tip5_permutation_coop(s, lane16, s_lut, s_rc);
    }
    state[lane16] = gl_canon(s);
}

inline unsigned grid_for(u64 count, int threads) { return (unsigned)((count + threads - 1) / threads); }

inline int launch_permute(u64 *d_states, u64 count, cudaStream_t st) {
    if (count == 0) return 0;
    TF21_LAUNCH(tip5_permute_kernel, grid_for(count, kTip5Threads), kTip5Threads, 0, st, d_states, count);
    return 0;
}

inline int launch_hash10(const u64 *d_in, u64 count, u64 *d_out, cudaStream_t st) {
    if (count == 0) return 0;
    TF21_LAUNCH_NAMED("tip5_hash10_kernel", tip5_hash10_kernel<false>, grid_for(count, kTip5Threads), kTip5Threads, 0,
                      st, d_in, count, d_out, (u64 *)nullptr);
    return 0;
}

// leaf level of a Merkle build: hash the pairs of d_leafs into d_out and copy the leaves to d_copy
inline int launch_hash10_copy(const u64 *d_in, u64 count, u64 *d_out, u64 *d_copy, cudaStream_t st) {
    if (count == 0) return 0;
    TF21_LAUNCH_NAMED("tip5_hash10_kernel", tip5_hash10_kernel<true>, grid_for(count, kTip5Threads), kTip5Threads, 0,
                      st, d_in, count, d_out, d_copy);
    return 0;
}

inline int launch_hash_rows(const u64 *d_data, u64 row_len, u64 n_rows, u64 row_stride, u64 elem_stride,
                            u64 *d_out, cudaStream_t st) {
    if (n_rows == 0) return 0;
    if (n_rows <= kMerkleCoopCnt) {  // latency bound: 16 lanes per row
        TF21_LAUNCH(tip5_hash_rows_coop_kernel, grid_for(16 * n_rows, kCoopThreads), kCoopThreads, 0, st, d_data, row_len,
                    n_rows, row_stride, elem_stride, d_out);
        return 0;
    }
    TF21_LAUNCH(tip5_hash_rows_kernel, grid_for(n_rows, kTip5Threads), kTip5Threads, 0, st, d_data, row_len,
                n_rows, row_stride, elem_stride, d_out);
    return 0;
}

constexpr u32 kMerkleTailCnt = 128;    // levels with <= this many nodes are finished by one CTA

// fills nodes[1..n) given nodes[n..2n) (sequentially_fill_tree, merkle_tree.rs:216-222, level-batched)
#ifndef TF21_MERKLE_TOP_CNT
#define TF21_MERKLE_TOP_CNT 8192
#endif
// levels with <= this many nodes are finished by merkle_top_kernel (0: never); TF21_MERKLE_TOP_CNT in the environment
// overrides the default for A/B runs
inline u64 merkle_top_cnt() {
    static const u64 v = [] {
        const char *e = getenv("TF21_MERKLE_TOP_CNT");
        return e ? (u64)strtoull(e, nullptr, 10) : (u64)TF21_MERKLE_TOP_CNT;
    }();
    return v;
}

// one cooperative launch for the levels cnt, cnt/2, ..., 1; returns 1 when the device refuses a cooperative grid
// (the caller then takes the level-by-level path), < 0 never, TF21 error codes otherwise
inline int launch_merkle_top(u64 *d_nodes, u64 cnt, cudaStream_t st) {
    int dev = 0, sms = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || !coop || sms <= 0) {
        cudaGetLastError();
        return 1;
    }
    u32 *bar = nullptr;
    TF21_CUDA(cudaMallocAsync((void **)&bar, sizeof(u32), st));
    cudaMemsetAsync(bar, 0, sizeof(u32), st);
    // no more CTAs than the first level has 16-lane groups for (a 2^8-leaf tree does not need 148 CTAs at a barrier)
    u64 grid = (cnt + kMerkleTopThreads / 16 - 1) / (kMerkleTopThreads / 16);
    if (grid > (u64)sms) grid = (u64)sms;
    u32 first_cnt = (u32)cnt;
    void *args[] = {(void *)&d_nodes, (void *)&first_cnt, (void *)&bar};
    ProfRec pr;
    const bool prof = g_prof_enabled.load(std::memory_order_relaxed);
    if (prof) prof_begin("merkle_top_kernel", st, &pr);
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)merkle_top_kernel, dim3((unsigned)grid),
                                                      dim3(kMerkleTopThreads), args, 0, st);
    if (prof) prof_end(st, &pr);
    cudaFreeAsync(bar, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        return 1;
    }
    if (e != cudaSuccess) return cuda_fail(e, "merkle_top_kernel", __LINE__);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

inline int launch_merkle_levels(u64 *d_nodes, u64 n_leafs, cudaStream_t st) {
    u64 cnt = n_leafs / 2;
    const u64 top = merkle_top_cnt();
    while (cnt > kMerkleTailCnt) {
        if (cnt <= top) {
            const int rc = launch_merkle_top(d_nodes, cnt, st);
            if (rc == 0) return 0;
            if (rc != 1) return rc;
        }
        if (cnt > kMerkleCoopCnt) {
            TF21_TRY(launch_hash10(d_nodes + 10 * cnt, cnt, d_nodes + 5 * cnt, st));
        } else {
            TF21_LAUNCH(tip5_hash10_coop_kernel, grid_for(16 * cnt, kCoopThreads), kCoopThreads, 0, st,
                        d_nodes + 10 * cnt, cnt, d_nodes + 5 * cnt);
        }
        cnt >>= 1;
    }
    if (cnt >= 1) {
        TF21_LAUNCH(merkle_tail_kernel, 1, kMerkleTailThreads, 0, st, d_nodes, (u32)cnt);
    }
    return 0;
}

inline int check_leaf_count(u64 n) {
    if (n == 0) return TF21_E_TOO_FEW_LEAFS;                    // merkle_tree.rs:394-396
    if (n & (n - 1)) return TF21_E_INCORRECT_NUMBER_OF_LEAFS;   // merkle_tree.rs:398-401
    if (n > (1ull << 40)) return TF21_E_ALLOC;                  // TreeTooHigh analogue
    return 0;
}

inline int merkle_build_dev(const u64 *d_leafs, u64 n, u64 *d_nodes, cudaStream_t st) {
    TF21_TRY(check_leaf_count(n));
    u64 words = 5 * n;
    if (n < 2 * (u64)kMerkleTailCnt) {  // small trees: copy, then the single-CTA tail
        unsigned grid = (unsigned)std::min<u64>((words + 255) / 256, 148 * 16);
        TF21_LAUNCH(merkle_init_kernel, grid, 256, 0, st, d_leafs, words, d_nodes);
        return launch_merkle_levels(d_nodes, n, st);
    }
    // nodes[0] = 0; the leaf level copies the leaves into nodes[n..2n) while hashing them (one read of the leaves)
    TF21_CUDA(cudaMemsetAsync(d_nodes, 0, 5 * sizeof(u64), st));
    TF21_TRY(launch_hash10_copy(d_leafs, n / 2, d_nodes + 5 * (n / 2), d_nodes + words, st));
    return launch_merkle_levels(d_nodes, n / 2, st);
}

}  // namespace tf21
