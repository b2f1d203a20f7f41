// host_stage.cuh -- pageable host slices through a pinned ring.
//
// The reference's callers hand over plain `Vec<BFieldElement>` / `&mut [XFieldElement]` slices
// (twenty-first/src/math/ntt.rs:67,109), i.e. PAGEABLE memory.  cudaMemcpyAsync on pageable memory is staged by the
// driver on the calling thread, one direction at a time, so the two PCIe directions never overlap.  Here a process-wide
// ring of pinned slots sits between the caller's slice and the device: copy-in threads fill slots and issue the H2D
// piece, copy-out threads drain the slots the D2H pieces landed in, and the calling thread only launches kernels --
// both PCIe directions, both host copies and the kernels run concurrently.  Pinned (or registered / managed) slices
// skip all of this (host_ntt in tf21.cu copies from them directly).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>

#include "runtime.cuh"

namespace tf21 {

constexpr size_t kStageSlotBytes = 4ull << 20;
constexpr int kStageSlots = 16;  // per direction

struct StageRing {
    char *in[kStageSlots] = {};
    char *out[kStageSlots] = {};
    cudaEvent_t in_done[kStageSlots] = {};
    cudaEvent_t out_done[kStageSlots] = {};
    bool ready = false;
    std::mutex busy;  // one staged call per device at a time (they would share the PCIe link anyway)
};

static std::mutex g_stage_mutex;                 // guards the map below
static std::map<int, StageRing> g_stage_rings;   // by device, guarded by g_stage_mutex

inline void stage_ring_free_all() {  // tf21_shutdown
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto &kv : g_stage_rings) {
        cudaSetDevice(kv.first);
        for (int i = 0; i < kStageSlots; i++) {
            if (kv.second.in[i]) cudaFreeHost(kv.second.in[i]);
            if (kv.second.out[i]) cudaFreeHost(kv.second.out[i]);
            if (kv.second.in_done[i]) cudaEventDestroy(kv.second.in_done[i]);
            if (kv.second.out_done[i]) cudaEventDestroy(kv.second.out_done[i]);
        }
    }
    g_stage_rings.clear();
    if (prev >= 0) cudaSetDevice(prev);
}

inline int stage_ring_get(int dev, StageRing **out) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    StageRing &r = g_stage_rings[dev];
    if (!r.ready) {
        for (int i = 0; i < kStageSlots; i++) {
            TF21_CUDA(cudaHostAlloc((void **)&r.in[i], kStageSlotBytes, cudaHostAllocDefault));
            TF21_CUDA(cudaHostAlloc((void **)&r.out[i], kStageSlotBytes, cudaHostAllocDefault));
            TF21_CUDA(cudaEventCreateWithFlags(&r.in_done[i], cudaEventDisableTiming));
            TF21_CUDA(cudaEventCreateWithFlags(&r.out_done[i], cudaEventDisableTiming));
        }
        r.ready = true;
    }
    *out = &r;
    return 0;
}

// Host copy between a caller's slice and a ring slot with non-temporal stores: the destination is not read again by
// this core (the DMA engine or the caller's later code reads it), so the read-for-ownership of every destination line
// and the cache pollution of a plain memcpy are saved.  glibc only switches to this form above its own threshold
// (several MiB per call); the ring moves 4 MiB pieces.
__attribute__((target("avx2"))) static void stage_copy_nt_avx2(char *dst, const char *src, size_t bytes) {
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head > bytes) head = bytes;
    if (head) memcpy(dst, src, head);
    dst += head, src += head, bytes -= head;
    const size_t body = bytes & ~(size_t)127;
    for (size_t o = 0; o < body; o += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + o));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(src + o + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + o + 64));
        const __m256i d = _mm256_loadu_si256((const __m256i *)(src + o + 96));
        _mm256_stream_si256((__m256i *)(dst + o), a);
        _mm256_stream_si256((__m256i *)(dst + o + 32), b);
        _mm256_stream_si256((__m256i *)(dst + o + 64), c);
        _mm256_stream_si256((__m256i *)(dst + o + 96), d);
    }
    _mm_sfence();
    if (bytes - body) memcpy(dst + body, src + body, bytes - body);
}
inline bool stage_copy_use_nt() {
    static const bool use = [] {
        const char *e = getenv("TF21_STAGE_NT");
        if (e && e[0] == '0') return false;
        return (bool)__builtin_cpu_supports("avx2");
    }();
    return use;
}
inline void stage_copy(void *dst, const void *src, size_t bytes) {
    if (bytes >= 4096 && stage_copy_use_nt()) stage_copy_nt_avx2((char *)dst, (const char *)src, bytes);
    else memcpy(dst, src, bytes);
}

// true for plain malloc / Vec memory (not pinned, not registered, not managed)
inline bool host_ptr_is_pageable(const void *p) {
    if (const char *e = getenv("TF21_NO_STAGE_RING"))
        if (e[0] == '1') return false;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// bytes per piece (<= slot size); TF21_STAGE_PIECE_KB for A/B runs
inline size_t stage_piece_bytes() {
    if (const char *e = getenv("TF21_STAGE_PIECE_KB")) {
        const size_t v = (size_t)strtoull(e, nullptr, 10) << 10;
        if (v >= (64u << 10) && v <= kStageSlotBytes) return v & ~(size_t)4095;
    }
    return kStageSlotBytes;
}

inline int stage_threads() {
    if (const char *e = getenv("TF21_STAGE_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 32) return v;
    }
    // measured on the 16-core bench host (tools/e2e_pageable.py): 2 / 4 / 6 / 8 threads per direction move 14 / 22 / 25 / 25 GB/s
    const unsigned hw = std::thread::hardware_concurrency();
    int t = (int)(hw / 2);
    return t < 2 ? 2 : (t > 8 ? 8 : t);
}

// One pass of  host chunk -> device buffer -> `compute` -> device buffer -> host chunk  over `n_chunks` chunks of a
// caller's pageable slice.  Chunk c occupies device buffer / stream c % n_lanes; `compute(c, lane)` launches the
// kernels of chunk c on streams[lane] (it runs on the calling thread, after every H2D piece of the chunk has been
// issued on that stream).
struct StagedPass {
    int dev = 0;
    StageRing *ring = nullptr;
    const char *host_in = nullptr;      // the caller's source slice (nullptr: no H2D leg)
    char *host_out = nullptr;           // the caller's destination slice (nullptr: no D2H leg)
    std::vector<size_t> chunk_off;      // n_chunks + 1 byte offsets into `host`
    int n_lanes = 0;
    cudaStream_t *streams = nullptr;
    char **dev_bufs = nullptr;
    std::function<int(size_t, int)> compute;

    // ---- shared state ----
    std::mutex m;
    std::condition_variable cv;
    int rc = 0;                             // first failure
    std::vector<size_t> piece_chunk, piece_off, piece_len;  // global piece list (same split for both directions)
    std::vector<size_t> chunk_first_piece;  // n_chunks + 1
    std::atomic<size_t> next_in{0};
    std::vector<size_t> issued_in_chunk;    // H2D pieces issued per chunk
    size_t in_slot_uses[kStageSlots] = {};  // pieces issued through each in-slot
    size_t d2h_enqueued_chunks = 0;
    bool out_free[kStageSlots];
    struct OutItem { size_t piece; int slot; };
    std::deque<OutItem> out_queue;
    bool out_closed = false;

    void fail(int code) {
        std::lock_guard<std::mutex> lock(m);
        if (rc == 0) rc = code;
        cv.notify_all();
    }

    void build_pieces() {
        const size_t n_chunks = chunk_off.size() - 1;
        chunk_first_piece.assign(n_chunks + 1, 0);
        for (size_t c = 0; c < n_chunks; c++) {
            chunk_first_piece[c] = piece_chunk.size();
            const size_t bytes = chunk_off[c + 1] - chunk_off[c];
            const size_t piece = stage_piece_bytes();
            for (size_t o = 0; o < bytes; o += piece) {
                piece_chunk.push_back(c);
                piece_off.push_back(o);
                piece_len.push_back(bytes - o < piece ? bytes - o : piece);
            }
        }
        chunk_first_piece[n_chunks] = piece_chunk.size();
        issued_in_chunk.assign(n_chunks, 0);
        for (int i = 0; i < kStageSlots; i++) out_free[i] = true;
    }

    void in_worker() {
        cudaSetDevice(dev);
        for (;;) {
            const size_t i = next_in.fetch_add(1);
            if (i >= piece_chunk.size()) return;
            const size_t c = piece_chunk[i];
            const int slot = (int)(i % kStageSlots);
            const int lane = (int)(c % (size_t)n_lanes);
            {
                // the slot's previous piece must have been issued (its event recorded), and the device buffer's previous
                // tenant (chunk c - n_lanes) must have all of its D2H pieces enqueued on the same stream
                std::unique_lock<std::mutex> lock(m);
                cv.wait(lock, [&] {
                    return rc != 0 || (in_slot_uses[slot] == i / kStageSlots &&
                                       (c < (size_t)n_lanes || d2h_enqueued_chunks + (size_t)n_lanes > c));
                });
                if (rc != 0) return;
            }
            if (i >= (size_t)kStageSlots && cudaEventSynchronize(ring->in_done[slot]) != cudaSuccess)
                return fail(cuda_fail(cudaGetLastError(), "stage: in-slot wait", __LINE__));
            stage_copy(ring->in[slot], host_in + chunk_off[c] + piece_off[i], piece_len[i]);
            if (cudaMemcpyAsync(dev_bufs[lane] + piece_off[i], ring->in[slot], piece_len[i], cudaMemcpyHostToDevice,
                                streams[lane]) != cudaSuccess ||
                cudaEventRecord(ring->in_done[slot], streams[lane]) != cudaSuccess)
                return fail(cuda_fail(cudaGetLastError(), "stage: H2D piece", __LINE__));
            {
                std::lock_guard<std::mutex> lock(m);
                in_slot_uses[slot]++;
                issued_in_chunk[c]++;
            }
            cv.notify_all();
        }
    }

    void out_worker() {
        cudaSetDevice(dev);
        for (;;) {
            OutItem it;
            {
                std::unique_lock<std::mutex> lock(m);
                cv.wait(lock, [&] { return rc != 0 || !out_queue.empty() || out_closed; });
                if (rc != 0) return;
                if (out_queue.empty()) return;  // closed and drained
                it = out_queue.front();
                out_queue.pop_front();
            }
            if (cudaEventSynchronize(ring->out_done[it.slot]) != cudaSuccess)
                return fail(cuda_fail(cudaGetLastError(), "stage: out-slot wait", __LINE__));
            stage_copy(host_out + chunk_off[piece_chunk[it.piece]] + piece_off[it.piece], ring->out[it.slot], piece_len[it.piece]);
            {
                std::lock_guard<std::mutex> lock(m);
                out_free[it.slot] = true;
            }
            cv.notify_all();
        }
    }

    int run() {
        build_pieces();
        const size_t n_chunks = chunk_off.size() - 1;
        const int nt = stage_threads();
        std::vector<std::thread> workers;
        try {  // nothing may unwind across the C ABI: a thread the OS refuses becomes an error code
            for (int t = 0; t < nt && host_in; t++) workers.emplace_back([this] { in_worker(); });
            for (int t = 0; t < nt && host_out; t++) workers.emplace_back([this] { out_worker(); });
        } catch (...) {
            fail(TF21_E_ALLOC);
        }
        size_t out_seq = 0;
        for (size_t c = 0; c < n_chunks; c++) {
            const int lane = (int)(c % (size_t)n_lanes);
            const size_t p0 = chunk_first_piece[c], p1 = chunk_first_piece[c + 1];
            {
                std::unique_lock<std::mutex> lock(m);
                cv.wait(lock, [&] { return rc != 0 || !host_in || issued_in_chunk[c] == p1 - p0; });
                if (rc != 0) break;
            }
            const int crc = compute(c, lane);
            if (crc != 0) {
                fail(crc);
                break;
            }
            bool ok = true;
            for (size_t i = p0; i < p1 && ok && host_out; i++, out_seq++) {
                const int slot = (int)(out_seq % kStageSlots);
                {
                    std::unique_lock<std::mutex> lock(m);
                    cv.wait(lock, [&] { return rc != 0 || out_free[slot]; });
                    if (rc != 0) {
                        ok = false;
                        break;
                    }
                    out_free[slot] = false;
                }
                if (cudaMemcpyAsync(ring->out[slot], dev_bufs[lane] + piece_off[i], piece_len[i], cudaMemcpyDeviceToHost,
                                    streams[lane]) != cudaSuccess ||
                    cudaEventRecord(ring->out_done[slot], streams[lane]) != cudaSuccess) {
                    fail(cuda_fail(cudaGetLastError(), "stage: D2H piece", __LINE__));
                    ok = false;
                    break;
                }
                {
                    std::lock_guard<std::mutex> lock(m);
                    out_queue.push_back(OutItem{i, slot});
                }
                cv.notify_all();
            }
            if (!ok) break;
            {
                std::lock_guard<std::mutex> lock(m);
                d2h_enqueued_chunks = c + 1;
            }
            cv.notify_all();
        }
        {
            std::lock_guard<std::mutex> lock(m);
            out_closed = true;
        }
        cv.notify_all();
        for (auto &w : workers) w.join();
        return rc;
    }
};

// ---- cudaMemcpy replacements for the host entry points: a large pageable slice goes through the ring --------------
constexpr size_t kStageMinBytes = 8ull << 20;

// dst[0, bytes) = src[0, bytes) with `stage_threads()` host threads (the leaf half of a Merkle node array is a copy of
// the caller's own leaves: it never crosses PCIe)
inline void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    const int nt = bytes < (4ull << 20) ? 1 : stage_threads();
    if (nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    size_t done = 0;
    try {
        for (int t = 0; t + 1 < nt && done + per < bytes; t++, done += per) {
            const size_t o = done;
            th.emplace_back([=] { stage_copy((char *)dst + o, (const char *)src + o, per); });
        }
    } catch (...) {  // fewer helpers than asked for: this thread copies what is left
    }
    stage_copy((char *)dst + done, (const char *)src + done, bytes - done);
    for (auto &x : th) x.join();
}

// H2D on `st`; on return the source slice may be reused (like cudaMemcpyAsync from pageable memory), the copy itself is
// ordered on `st`
inline int copy_h2d(void *d, const void *h, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    if (bytes < kStageMinBytes || !host_ptr_is_pageable(h)) {
        TF21_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    StagedPass sp;
    TF21_TRY(current_device(&sp.dev));
    TF21_TRY(stage_ring_get(sp.dev, &sp.ring));
    std::lock_guard<std::mutex> stage_lock(sp.ring->busy);
    char *dev = (char *)d;
    sp.host_in = (const char *)h;
    sp.chunk_off = {0, bytes};
    sp.n_lanes = 1;
    sp.streams = &st;
    sp.dev_bufs = &dev;
    sp.compute = [](size_t, int) { return 0; };
    TF21_TRY(sp.run());
    // the ring slots are reused by the next staged call: their H2D pieces must have left them
    TF21_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// D2H, complete on return (like cudaMemcpy); the device data must be ready in stream order on `st`
inline int copy_d2h(void *h, const void *d, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    if (bytes < kStageMinBytes || !host_ptr_is_pageable(h)) {
        TF21_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
        TF21_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    StagedPass sp;
    TF21_TRY(current_device(&sp.dev));
    TF21_TRY(stage_ring_get(sp.dev, &sp.ring));
    std::lock_guard<std::mutex> stage_lock(sp.ring->busy);
    char *dev = (char *)const_cast<void *>(d);
    sp.host_out = (char *)h;
    sp.chunk_off = {0, bytes};
    sp.n_lanes = 1;
    sp.streams = &st;
    sp.dev_bufs = &dev;
    sp.compute = [](size_t, int) { return 0; };
    return sp.run();
}

}  // namespace tf21
