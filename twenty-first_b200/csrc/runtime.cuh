// runtime.cuh -- host runtime of libtf21: error plumbing, per-device table cache, launch counter.
//
// The reference caches its per-size twiddle and swap tables behind OnceLock statics
// (twenty-first/src/math/ntt.rs:71,113,166) and must stay callable from any rayon worker; the
// equivalent here is a mutex-guarded per-(device,size,direction) cache of device-resident tables,
// built on the host (64-bit modular arithmetic) and uploaded once.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/tf21.h"
#include "field.cuh"

namespace tf21 {

static thread_local char g_last_cuda_error[256] = "";
static std::atomic<uint64_t> g_launches{0};

inline int cuda_fail(cudaError_t e, const char *what, int line) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s (line %d): %s", what, line, cudaGetErrorString(e));
    return (e == cudaErrorMemoryAllocation) ? TF21_E_ALLOC : TF21_E_CUDA;
}

#define TF21_CUDA(call)                                                    \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) return tf21::cuda_fail(_e, #call, __LINE__); \
    } while (0)

#define TF21_TRY(call)        \
    do {                      \
        int _rc = (call);     \
        if (_rc != 0) return _rc; \
    } while (0)

// Optional per-launch timing (tf21_profile_enable): CUDA events recorded on the launching stream
// around every kernel, read back by tf21_profile_read.  Off by default; bench.py turns it on for
// its roofline pass only.
struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
};
static std::atomic<bool> g_prof_enabled{false};
static std::mutex g_prof_mutex;
static std::vector<ProfRec> g_prof;

inline void prof_begin(const char *name, cudaStream_t st, ProfRec *r) {
    r->name = name;
    cudaEventCreate(&r->a);
    cudaEventCreate(&r->b);
    cudaEventRecord(r->a, st);
}
inline void prof_end(cudaStream_t st, ProfRec *r) {
    cudaEventRecord(r->b, st);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof.push_back(*r);
}

// every kernel launch of the library goes through this so bench.py can report gpu_launches
#define TF21_LAUNCH_NAMED(name, kernel, grid, block, smem, stream, ...)          \
    do {                                                                        \
        tf21::ProfRec _pr;                                                      \
        const bool _prof = tf21::g_prof_enabled.load(std::memory_order_relaxed); \
        if (_prof) tf21::prof_begin((name), (stream), &_pr);                    \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
        if (_prof) tf21::prof_end((stream), &_pr);                              \
        tf21::g_launches.fetch_add(1, std::memory_order_relaxed);               \
        cudaError_t _e = cudaGetLastError();                                    \
        if (_e != cudaSuccess) return tf21::cuda_fail(_e, (name), __LINE__);    \
    } while (0)
#define TF21_LAUNCH(kernel, grid, block, smem, stream, ...) \
    TF21_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

inline unsigned ilog2_u64(uint64_t n) { return 63u - (unsigned)__builtin_clzll(n); }

// ---- device table cache ---------------------------------------------------------------------
struct DeviceTables {
    // omega_{2^l}^{+-e}, e < 2^(l-1), for l = 1..10, packed at offset 2^(l-1)-1; [0]=fwd, [1]=inv
    u64 *tw_small[2] = {nullptr, nullptr};
    // split tables for omega_B^{+-e}: key (log2 B, inverse)
    struct Split {
        u64 *lo = nullptr, *hi = nullptr;
        unsigned h = 0;
    };
    std::map<std::pair<unsigned, int>, Split> split;
    // coset scale tables: key (g, c0, h, n_lo, n_hi)
    std::map<std::tuple<u64, u64, unsigned, u64, u64>, Split> scale;
    std::vector<void *> owned;
    bool constants_ready = false;
    int sm_count = 0;
    size_t smem_optin = 0;
};

static std::mutex g_mutex;
static std::map<int, DeviceTables> g_devices;

inline int current_device(int *dev) {
    TF21_CUDA(cudaGetDevice(dev));
    return 0;
}

// Tables are read by kernels on caller streams, which may be cudaStreamNonBlocking (host_ntt's own streams, torch
// side streams) and are then not ordered against the legacy stream.  A cudaMemcpy from pageable memory may return
// once the data is staged, before the DMA to the device has finished -- so the copy is drained here, before the
// pointer is published in a cache.
inline int upload(DeviceTables &t, const std::vector<u64> &host, u64 **out) {
    void *d = nullptr;
    TF21_CUDA(cudaMalloc(&d, host.size() * sizeof(u64)));
    cudaError_t e = cudaMemcpy(d, host.data(), host.size() * sizeof(u64), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) {
        cudaFree(d);
        return cuda_fail(e, "upload", __LINE__);
    }
    t.owned.push_back(d);
    *out = (u64 *)d;
    return 0;
}

// free one table of `owned` (used by bounded caches); in-flight kernels may still read it: drain the device first
inline void release_owned(DeviceTables &t, void *p) {
    for (auto it = t.owned.begin(); it != t.owned.end(); ++it)
        if (*it == p) {
            t.owned.erase(it);
            break;
        }
    cudaFree(p);
}

// omega_B^{e}, B = 2^lb, as lo[e & (2^h - 1)] * hi[e >> h]
inline int get_split_tables(DeviceTables &t, unsigned lb, int inverse, DeviceTables::Split *out) {
    auto key = std::make_pair(lb, inverse);
    auto it = t.split.find(key);
    if (it != t.split.end()) {
        *out = it->second;
        return 0;
    }
    DeviceTables::Split s;
    s.h = (lb + 1) / 2;
    u64 w = hgl_root_of_unity(lb);
    if (inverse) w = hgl_inv(w);
    u64 n_lo = 1ull << s.h, n_hi = 1ull << (lb - s.h);
    std::vector<u64> lo(n_lo), hi(n_hi);
    u64 acc = 1;
    for (u64 e = 0; e < n_lo; e++) {
        lo[e] = acc;
        acc = hgl_mul(acc, w);
    }
    u64 step = acc;  // w^(2^h)
    acc = 1;
    for (u64 e = 0; e < n_hi; e++) {
        hi[e] = acc;
        acc = hgl_mul(acc, step);
    }
    TF21_TRY(upload(t, lo, &s.lo));
    TF21_TRY(upload(t, hi, &s.hi));
    t.split[key] = s;
    *out = s;
    return 0;
}

// c0 * g^i for i < count, as lo[i & (2^h-1)] * hi[i >> h]   (c0 folded into lo)
inline int get_scale_tables(DeviceTables &t, u64 g, u64 c0, u64 count, DeviceTables::Split *out) {
    unsigned lc = count <= 1 ? 0 : ilog2_u64(count - 1) + 1;
    unsigned h = (lc + 1) / 2;
    u64 n_lo = 1ull << h, n_hi = 1ull << (lc - h);
    auto key = std::make_tuple(g, c0, h, n_lo, n_hi);
    auto it = t.scale.find(key);
    if (it != t.scale.end()) {
        *out = it->second;
        return 0;
    }
    if (t.scale.size() >= 64) {  // bounded cache: drop every entry and free its tables (kernels in flight on
        cudaDeviceSynchronize();  // other streams may still read them, so the device is drained first)
        for (auto &kv : t.scale) {
            release_owned(t, kv.second.lo);
            release_owned(t, kv.second.hi);
        }
        t.scale.clear();
    }
    DeviceTables::Split s;
    s.h = h;
    std::vector<u64> lo(n_lo), hi(n_hi);
    u64 acc = c0 % GL_P;
    u64 pw = 1;
    for (u64 e = 0; e < n_lo; e++) {
        lo[e] = acc;
        acc = hgl_mul(acc, g);
        pw = hgl_mul(pw, g);
    }
    acc = 1;
    for (u64 e = 0; e < n_hi; e++) {
        hi[e] = acc;
        acc = hgl_mul(acc, pw);
    }
    TF21_TRY(upload(t, lo, &s.lo));
    TF21_TRY(upload(t, hi, &s.hi));
    t.scale[key] = s;
    *out = s;
    return 0;
}

}  // namespace tf21
