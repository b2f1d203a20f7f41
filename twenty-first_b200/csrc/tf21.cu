// tf21.cu -- the C ABI of libtf21 (include/tf21.h): argument checking, device state, host<->device
// staging, and dispatch into the sm_100a kernels.  Single translation unit (the kernels live in the
// included .cuh files) so the __constant__ tables are shared without relocatable device code.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include "ntt_fast.cuh"
#include "tip5_kernels.cuh"
#include "host_stage.cuh"

using namespace tf21;

namespace tf21 {

// per-device state, created on first use of a device
static int get_tables(DeviceTables **out, int *dev_out = nullptr) {
    int dev;
    TF21_TRY(current_device(&dev));
    if (dev_out) *dev_out = dev;
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceTables &t = g_devices[dev];
    if (!t.constants_ready) {
        cudaDeviceProp prop;
        TF21_CUDA(cudaGetDeviceProperties(&prop, dev));
        t.sm_count = prop.multiProcessorCount;
        t.smem_optin = prop.sharedMemPerBlockOptin;
        {   // keep stream-ordered scratch (cudaMallocAsync) cached across synchronisations: the default
            // release threshold of 0 returns the memory to the OS at every sync and re-maps it per call
            cudaMemPool_t pool;
            TF21_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            uint64_t thr = ~0ull;
            TF21_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        }
        TF21_TRY(upload_tip5_constants());
        {
            u64 w64[64];
            const u64 w = hgl_root_of_unity(6);
            for (int k = 0; k < 64; k++) w64[k] = hgl_pow(w, k);
            TF21_CUDA(cudaMemcpyToSymbol(c_w64, w64, sizeof(w64)));
            TF21_CUDA(cudaStreamSynchronize(nullptr));
        }
        for (int inv = 0; inv < 2; inv++) {
            std::vector<u64> tw((1u << kNttMaxLogPass) - 1);
            for (u32 l = 1; l <= kNttMaxLogPass; l++) {
                u64 w = hgl_root_of_unity(l);
                if (inv) w = hgl_inv(w);
                u64 acc = 1;
                u64 *dst = tw.data() + ((1u << l) >> 1) - 1;
                for (u32 e = 0; e < (1u << l) / 2; e++) {
                    dst[e] = acc;
                    acc = hgl_mul(acc, w);
                }
            }
            TF21_TRY(upload(t, tw, &t.tw_small[inv]));
        }
        TF21_CUDA(cudaFuncSetAttribute(ntt_col_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)t.smem_optin));
        TF21_CUDA(cudaFuncSetAttribute(ntt_row_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)t.smem_optin));
#define TF21_FAST_SMEM(K) TF21_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem))
        TF21_FAST_SMEM((ntt1024_col_kernel<false, false>));
        TF21_FAST_SMEM((ntt1024_col_kernel<false, true>));
        TF21_FAST_SMEM((ntt1024_col_kernel<true, false>));
        TF21_FAST_SMEM((ntt1024_col_kernel<true, true>));
        TF21_FAST_SMEM((ntt1024_row_kernel<false, 1, false>));
        TF21_FAST_SMEM((ntt1024_row_kernel<false, 1, true>));
        TF21_FAST_SMEM((ntt1024_row_kernel<true, 1, false>));
        TF21_FAST_SMEM((ntt1024_row_kernel<true, 1, true>));
        TF21_FAST_SMEM((ntt1024_row_kernel<false, 3, false>));
        TF21_FAST_SMEM((ntt1024_row_kernel<false, 3, true>));
        TF21_FAST_SMEM((ntt1024_row_kernel<true, 3, false>));
        TF21_FAST_SMEM((ntt1024_row_kernel<true, 3, true>));
#undef TF21_FAST_SMEM
#define TF21_SMALL_N_SMEM(K_)                                                                                          \
    TF21_CUDA(cudaFuncSetAttribute(ntt_small_n_kernel<false, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem)); \
    TF21_CUDA(cudaFuncSetAttribute(ntt_small_n_kernel<true, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));  \
    TF21_CUDA(cudaFuncSetAttribute(ntt_col_n_kernel<false, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));   \
    TF21_CUDA(cudaFuncSetAttribute(ntt_col_n_kernel<true, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
        TF21_SMALL_N_SMEM(1) TF21_SMALL_N_SMEM(2) TF21_SMALL_N_SMEM(3) TF21_SMALL_N_SMEM(4) TF21_SMALL_N_SMEM(5)
        TF21_SMALL_N_SMEM(6) TF21_SMALL_N_SMEM(7) TF21_SMALL_N_SMEM(8) TF21_SMALL_N_SMEM(9)
#undef TF21_SMALL_N_SMEM
#define TF21_TMA_SMEM(K) TF21_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColTmaSmem))
        TF21_TMA_SMEM((ntt1024_col_tma_kernel<false, false>));
        TF21_TMA_SMEM((ntt1024_col_tma_kernel<false, true>));
        TF21_TMA_SMEM((ntt1024_col_tma_kernel<true, false>));
        TF21_TMA_SMEM((ntt1024_col_tma_kernel<true, true>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<false, 1, true>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<true, 1, true>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<false, 3, true>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<true, 3, true>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<false, 1, false>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<true, 1, false>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<false, 3, false>));
        TF21_TMA_SMEM((ntt1024_row_tma_kernel<true, 3, false>));
#undef TF21_TMA_SMEM
        TF21_CUDA(cudaFuncSetAttribute(tma_tile_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTmaTileWords * 8 + 1024 + 16)));
        TF21_CUDA(cudaFuncSetAttribute(ntt1024_single_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
        TF21_CUDA(cudaFuncSetAttribute(ntt1024_single_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
        TF21_CUDA(cudaFuncSetAttribute(ntt1024_single_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
        TF21_CUDA(cudaFuncSetAttribute(ntt1024_single_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
        t.constants_ready = true;
    }
    *out = &t;
    return 0;
}

// ---- NCCL, bound at run time -------------------------------------------------------------------------
// The only exchange on the path is the Merkle tree cap (one 40-byte root per shard).  NCCL is
// dlopen'ed (libnccl.so.2: torch's bundled copy if the process already holds it, else the system one) so that
// single-GPU users need no NCCL at all; the handful of prototypes below are the stable C API of nccl.h.
typedef struct ncclComm *tf21_ncclComm_t;
struct NcclApi {
    void *handle = nullptr;
    int (*CommInitAll)(tf21_ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(tf21_ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, tf21_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool tried = false;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;
static std::map<int, std::vector<tf21_ncclComm_t>> g_nccl_comms;  // by device count: communicators of devices 0..n-1
static int g_n_devices = 0;                                        // tf21_init: devices the sharded entry points use (0 = all)
constexpr int kNcclUint64 = 5;                                     // ncclUint64, nccl.h

static int nccl_fail(int rc, const char *what) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "NCCL %s: %s", what,
             (g_nccl.GetErrorString && rc >= 0) ? g_nccl.GetErrorString(rc) : "library not available");
    return TF21_E_NCCL;
}

// caller holds g_nccl_mutex
static bool nccl_load() {
    if (g_nccl.tried) return g_nccl.handle != nullptr;
    g_nccl.tried = true;
    if (getenv("TF21_NO_NCCL")) return false;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(h, "ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        dlclose(h);
        return false;
    }
    g_nccl.handle = h;
    return true;
}

// communicators over devices 0..n-1 of this process (ncclCommInitAll), created once per n
static int nccl_comms(int n, tf21_ncclComm_t **out) {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (!nccl_load()) return nccl_fail(-1, "dlopen(libnccl.so.2)");
    auto it = g_nccl_comms.find(n);
    if (it == g_nccl_comms.end()) {
        std::vector<tf21_ncclComm_t> comms(n);
        std::vector<int> devs(n);
        for (int i = 0; i < n; i++) devs[i] = i;
        int prev = 0;
        cudaGetDevice(&prev);
        const int rc = g_nccl.CommInitAll(comms.data(), n, devs.data());
        cudaSetDevice(prev);
        if (rc != 0) return nccl_fail(rc, "ncclCommInitAll");
        it = g_nccl_comms.emplace(n, std::move(comms)).first;
    }
    *out = it->second.data();
    return 0;
}

static int check_ntt_len(u64 n, u32 width) {
    if (width != 1 && width != 3) return TF21_E_BAD_ARG;
    if (n > 0xffffffffull) return TF21_E_LEN_TOO_LARGE;          // ntt.rs:135-136
    if (n != 0 && (n & (n - 1)) != 0) return TF21_E_LEN_NOT_POW2; // ntt.rs:137
    if (n > (1ull << 30)) return TF21_E_LEN_TOO_LARGE;           // device implementation limit
    return 0;
}

// the Tip5 kernels read hash inputs / states with 16-byte loads (10 or 16 words per item keep that alignment from
// an aligned base); a pointer that is only 8-byte aligned is refused instead of faulting on the device
static inline bool misaligned16(const void *p) { return ((uintptr_t)p & 15) != 0; }

// stream-ordered scratch
struct Scratch {
    u64 *p = nullptr;
    cudaStream_t st;
    explicit Scratch(cudaStream_t s) : st(s) {}
    int alloc(u64 words) {
        if (words == 0) return 0;
        TF21_CUDA(cudaMallocAsync((void **)&p, words * sizeof(u64), st));
        return 0;
    }
    ~Scratch() {
        if (p) cudaFreeAsync(p, st);
    }
};

// table builders take the mutex themselves
static int split_locked(DeviceTables &t, u64 g, u64 c0, u64 count, ScaleTab *out) {
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceTables::Split s;
    TF21_TRY(get_scale_tables(t, g, c0, count, &s));
    *out = ScaleTab{s.lo, s.hi, s.h};
    return 0;
}

static u64 l2_group_cols() {
    const char *e = getenv("TF21_L2_GROUP");
    return e ? (u64)strtoull(e, nullptr, 10) : 0;
}

static int ntt_run_locked(DeviceTables &t, const u64 *src, u64 n_in, u64 *dst, u64 n, u32 w, u64 batch, int inverse,
                          ScaleTab pre, ScaleTab post, u64 post_scalar, cudaStream_t st) {
    if (n <= 1) {  // identity transform (ntt.rs:178-181 returns early; len 1 has no stages)
        if (n == 1 && src != dst && n_in >= 1)
            TF21_CUDA(cudaMemcpyAsync(dst, src, batch * w * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        if (n == 1 && n_in == 0) TF21_CUDA(cudaMemsetAsync(dst, 0, batch * w * sizeof(u64), st));
        return 0;
    }
    int dev;
    TF21_TRY(current_device(&dev));
    // Column-group schedule (TF21_L2_GROUP = columns per group): the batch is cut into groups whose inter-pass scratch
    // (group * n * w * 8 bytes) fits the 126 MB L2, and the groups alternate between two side streams, so that the
    // tail wave of one group's pass overlaps the other group's kernels while the scratch written by the column
    // pass is still L2-resident when the row pass reads it.  Scratch shrinks from the batch to two groups.
    const u64 group = l2_group_cols();
    if (group > 0 && ilog2_u64(n) > kNttMaxLogPass && batch >= 4 * group) {
        cudaStream_t ss[2] = {nullptr, nullptr};
        cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
        int rc = 0;
        for (int i = 0; i < 3 && rc == 0; i++)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "cudaEventCreate", __LINE__);
        if (rc == 0 && cudaEventRecord(ev[2], st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaEventRecord", __LINE__);
        for (int i = 0; i < 2 && rc == 0; i++)
            if (cudaStreamCreateWithFlags(&ss[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamWaitEvent(ss[i], ev[2], 0) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "group stream", __LINE__);
        if (rc == 0) {
            Scratch s0(ss[0]), s1(ss[1]);
            rc = s0.alloc(n * w * group);
            if (rc == 0) rc = s1.alloc(n * w * group);
            for (u64 b0 = 0, g = 0; b0 < batch && rc == 0; b0 += group, g++) {
                const u64 cnt = batch - b0 < group ? batch - b0 : group;
                rc = ntt_run(t, dev, src + b0 * n_in * w, n_in, dst + b0 * n * w, n, w, cnt, inverse, pre, post, post_scalar,
                             (g & 1) ? s1.p : s0.p, ss[g & 1]);
            }
        }
        for (int i = 0; i < 2; i++) {
            if (ss[i] && ev[i] && cudaEventRecord(ev[i], ss[i]) == cudaSuccess) cudaStreamWaitEvent(st, ev[i], 0);
            if (ss[i]) cudaStreamDestroy(ss[i]);  // released once its work has drained
        }
        for (int i = 0; i < 3; i++)
            if (ev[i]) cudaEventDestroy(ev[i]);
        return rc;
    }
    Scratch scratch(st);
    if (ilog2_u64(n) > kNttMaxLogPass) TF21_TRY(scratch.alloc(n * w * batch));
    return ntt_run(t, dev, src, n_in, dst, n, w, batch, inverse, pre, post, post_scalar, scratch.p, st);
}

}  // namespace tf21

#define NO_SCALE (ScaleTab{nullptr, nullptr, 0})

extern "C" {

// tf21_init(n_devices), 0 = all visible.  Only the calling thread's current device is prepared here;
// the other devices (and the NCCL communicators of the sharded Merkle build) are set up on first use, so that a
// one-process-per-GPU launch (torchrun: every rank sees all eight GPUs) never opens contexts it will not use.
int tf21_init(int n_devices) {
    int visible = 0;
    TF21_CUDA(cudaGetDeviceCount(&visible));
    if (n_devices < 0 || n_devices > visible) return TF21_E_BAD_ARG;
    {
        std::lock_guard<std::mutex> lock(g_nccl_mutex);
        g_n_devices = n_devices;
    }
    DeviceTables *t;
    return get_tables(&t);
}

int tf21_set_device(int device) {
    TF21_CUDA(cudaSetDevice(device));
    DeviceTables *t;
    return get_tables(&t);
}

int tf21_device_count(void) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    return (g_n_devices > 0 && g_n_devices < visible) ? g_n_devices : visible;
}

int tf21_shutdown(void) {
    stage_ring_free_all();
    {
        std::lock_guard<std::mutex> nlock(g_nccl_mutex);
        for (auto &kv : g_nccl_comms)
            for (auto c : kv.second) g_nccl.CommDestroy(c);
        g_nccl_comms.clear();
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto &kv : g_devices) {
        cudaSetDevice(kv.first);
        cudaDeviceSynchronize();
        for (void *p : kv.second.owned) cudaFree(p);
    }
    g_devices.clear();
    g_fast_tables.clear();  // their device pointers lived in `owned` and are gone: never hand them out again
    g_small_n_tw.clear();
    g_col_n_tw1.clear();
    g_mid_tw1.clear();
    if (prev >= 0) cudaSetDevice(prev);
    return 0;
}

const char *tf21_strerror(int code) {
    switch (code) {
        case TF21_OK: return "ok";
        case TF21_E_LEN_NOT_POW2: return "slice length must be 0 or a power of two (ntt.rs:137)";
        case TF21_E_LEN_TOO_LARGE: return "slice should be no longer than u32::MAX (ntt.rs:135); device limit 2^30";
        case TF21_E_TOO_FEW_LEAFS: return "MerkleTreeError::TooFewLeafs";
        case TF21_E_INCORRECT_NUMBER_OF_LEAFS: return "MerkleTreeError::IncorrectNumberOfLeafs";
        case TF21_E_ORDER_LE_DEGREE:
            return "`Polynomial::fast_coset_evaluate` is currently limited to domains of order greater than the "
                   "degree of the polynomial.";
        case TF21_E_ALLOC: return "allocation failed (MerkleTreeError::TreeTooHigh)";
        case TF21_E_CUDA: return "CUDA error (see tf21_last_cuda_error)";
        case TF21_E_BAD_ARG: return "bad argument";
        case TF21_E_LEAF_INDEX_INVALID: return "MerkleTreeError::LeafIndexInvalid";
        case TF21_E_CAPACITY: return "output buffer too small";
        case TF21_E_DIVISION_BY_ZERO: return "divisor should be non-zero";
        case TF21_E_NCCL: return "NCCL error (see tf21_last_cuda_error)";
        default: return "unknown tf21 error";
    }
}

const char *tf21_last_cuda_error(void) { return g_last_cuda_error; }
uint64_t tf21_kernel_launch_count(void) { return g_launches.load(); }

int tf21_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto &r : g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    g_prof_enabled.store(on != 0);
    return 0;
}

// writes "kernel_name milliseconds\n" per recorded launch, in launch order; returns the number of
// bytes needed (call again with a larger buffer if it exceeds buflen)
int64_t tf21_profile_read(char *buf, uint64_t buflen) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    std::string out;
    for (auto &r : g_prof) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) return TF21_E_CUDA;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) return TF21_E_CUDA;
        char line[160];
        snprintf(line, sizeof(line), "%s %.6f\n", r.name, ms);
        out += line;
    }
    if (buf && buflen) {
        size_t ncopy = out.size() < buflen - 1 ? out.size() : buflen - 1;
        memcpy(buf, out.data(), ncopy);
        buf[ncopy] = 0;
    }
    return (int64_t)out.size() + 1;
}

int tf21_malloc(void **dptr, uint64_t bytes) {
    TF21_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
}
int tf21_free(void *dptr) {
    TF21_CUDA(cudaFree(dptr));
    return 0;
}
int tf21_memcpy_h2d(void *dst, const void *src, uint64_t bytes, tf21_stream_t stream) {
    TF21_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}
int tf21_memcpy_d2h(void *dst, const void *src, uint64_t bytes, tf21_stream_t stream) {
    TF21_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}
int tf21_stream_sync(tf21_stream_t stream) {
    TF21_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ---- diagnostics -----------------------------------------------------------------------------------
__global__ void selftest_field_kernel(int op, const u64 *a, const u64 *b, u64 *out, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 x = a[i], y = b[i], r = 0;
    switch (op) {
        case 0: r = gl_add(x, y); break;        // canonical operands
        case 1: r = gl_sub(x, y); break;        // canonical operands
        case 2: r = gl_mulc(x, y); break;       // any operands
        case 3: r = gl_canon(x); break;
        case 4: r = gl_canon(gl_add_weak(x, y)); break;  // x any, y canonical
        case 5: r = gl_canon(gl_reduce96(x, (u32)y)); break;
        case 6: r = gl_canon(gl_addl(x, y)); break;  // x any, y <= p
        case 7: r = gl_canonw(x); break;
        case 8: r = gl_canon(gl_sub(x, y)); break;   // x any, y < p
        case 9: r = gl_canon(gl_subl(x, y)); break;  // x any, y <= p (butterfly form)
#define TF21_SHL_CASE(S) case 100 + (S): r = gl_shlc<S>(x); break;
            TF21_SHL_CASE(3) TF21_SHL_CASE(6) TF21_SHL_CASE(9) TF21_SHL_CASE(12) TF21_SHL_CASE(15)
            TF21_SHL_CASE(18) TF21_SHL_CASE(21) TF21_SHL_CASE(24) TF21_SHL_CASE(27) TF21_SHL_CASE(30)
            TF21_SHL_CASE(33) TF21_SHL_CASE(36) TF21_SHL_CASE(39) TF21_SHL_CASE(42) TF21_SHL_CASE(45)
            TF21_SHL_CASE(48) TF21_SHL_CASE(51) TF21_SHL_CASE(54) TF21_SHL_CASE(57) TF21_SHL_CASE(60)
            TF21_SHL_CASE(63) TF21_SHL_CASE(66) TF21_SHL_CASE(69) TF21_SHL_CASE(72) TF21_SHL_CASE(75)
            TF21_SHL_CASE(78) TF21_SHL_CASE(81) TF21_SHL_CASE(84) TF21_SHL_CASE(87) TF21_SHL_CASE(90)
            TF21_SHL_CASE(93) TF21_SHL_CASE(1) TF21_SHL_CASE(31) TF21_SHL_CASE(95) TF21_SHL_CASE(65)
#undef TF21_SHL_CASE
    }
    out[i] = r;
}

int tf21_selftest_field_dev(int op, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, uint64_t n,
                            tf21_stream_t stream) {
    if (n == 0) return 0;
    TF21_LAUNCH(selftest_field_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, op, d_a, d_b, d_out, n);
    return 0;
}

// lands n_tiles tiles ([1024 rows][4 words] each, tile t = word-columns 4t..4t+3 of a [1024][inner_words] matrix) by
// TMA and returns the raw shared-memory images (4096 words per tile): pins the swizzle formula of tma.cuh
int tf21_selftest_tma_tile_dev(const uint64_t *d_matrix, uint64_t inner_words, uint64_t n_tiles, uint64_t *d_out,
                               tf21_stream_t stream) {
    if (!d_matrix || !d_out || n_tiles == 0 || n_tiles * kTmaTileCols > inner_words) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    CUtensorMap map;
    if (!tma_encode_tile_map(&map, d_matrix, inner_words, 1)) return TF21_E_BAD_ARG;
    TF21_LAUNCH(tma_tile_probe_kernel, (unsigned)n_tiles, 128, kTmaTileWords * 8 + 1024 + 16, (cudaStream_t)stream, map,
                d_out);
    return 0;
}

// ---- NTT -------------------------------------------------------------------------------------------
int tf21_ntt_dev(uint64_t *d_data, uint64_t n, uint32_t width, uint64_t batch, int inverse, tf21_stream_t stream) {
    TF21_TRY(check_ntt_len(n, width));
    if (n <= 1 || batch == 0) return 0;
    if (!d_data) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    u64 post_scalar = inverse ? hgl_inv(n % GL_P) : 0;  // unscale, ntt.rs:220-228
    return ntt_run_locked(*t, d_data, n, d_data, n, width, batch, inverse, NO_SCALE, NO_SCALE, post_scalar,
                          (cudaStream_t)stream);
}

// host staging helper: device buffer with H2D on construction side and D2H on demand
// (stream-ordered on the legacy stream the host entry points work on: the pool keeps the memory cached between calls,
// a cudaMalloc / cudaFree pair of 2 GB costs milliseconds and a device-wide synchronisation)
struct DevBuf {
    u64 *p = nullptr;
    int alloc(u64 words) {
        TF21_CUDA(cudaMallocAsync((void **)&p, (words ? words : 1) * sizeof(u64), nullptr));
        return 0;
    }
    ~DevBuf() {
        if (p) cudaFreeAsync(p, nullptr);
    }
};

// Host-slice NTT.  A batch is cut into chunks of whole arrays that go H2D -> kernels -> D2H on three
// streams, so the two PCIe directions and the compute overlap (the copies dominate: PCIe is ~100x
// slower than HBM).  Pinned host memory gets the full benefit; pageable memory is still correct.
static int host_ntt(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch, int inverse) {
    TF21_TRY(check_ntt_len(n, width));
    if (n <= 1 || batch == 0) return 0;
    if (!data) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    const u64 array_words = n * width;
    const u64 kChunkBytes = 64ull << 20;
    u64 chunk_arrays = kChunkBytes / (array_words * sizeof(u64));
    if (chunk_arrays < 1) chunk_arrays = 1;
    if (chunk_arrays > batch) chunk_arrays = batch;
    const u64 n_chunks = (batch + chunk_arrays - 1) / chunk_arrays;
    const int n_streams = (int)(n_chunks < 3 ? n_chunks : 3);
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    u64 *bufs[3] = {nullptr, nullptr, nullptr};
    int rc = 0;
    const bool pageable = host_ptr_is_pageable(data);
    auto cleanup = [&]() {
        for (int i = 0; i < n_streams; i++) {
            if (streams[i]) {
                if (bufs[i]) cudaFreeAsync(bufs[i], streams[i]);
                cudaStreamSynchronize(streams[i]);
                cudaStreamDestroy(streams[i]);
            }
        }
    };
    for (int i = 0; i < n_streams && rc == 0; i++) {
        if (cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaMallocAsync((void **)&bufs[i], chunk_arrays * array_words * sizeof(u64), streams[i]) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "host_ntt staging", __LINE__);
    }
    if (pageable && rc == 0) {
        // a plain Vec / malloc slice: through the pinned ring (host_stage.cuh), one staged call at a time
        StagedPass sp;
        rc = current_device(&sp.dev);
        if (rc == 0) rc = stage_ring_get(sp.dev, &sp.ring);
        if (rc == 0) {
            std::lock_guard<std::mutex> stage_lock(sp.ring->busy);
            sp.host_in = (const char *)data;
            sp.host_out = (char *)data;
            for (u64 c = 0; c <= n_chunks; c++) {
                const u64 a0 = c * chunk_arrays < batch ? c * chunk_arrays : batch;
                sp.chunk_off.push_back((size_t)(a0 * array_words * sizeof(u64)));
            }
            sp.n_lanes = n_streams;
            sp.streams = streams;
            sp.dev_bufs = (char **)bufs;
            sp.compute = [&](size_t c, int lane) {
                const u64 a0 = (u64)c * chunk_arrays;
                const u64 cnt = (a0 + chunk_arrays <= batch) ? chunk_arrays : batch - a0;
                return tf21_ntt_dev(bufs[lane], n, width, cnt, inverse, (tf21_stream_t)streams[lane]);
            };
            rc = sp.run();
        }
    }
    for (u64 c = 0; c < n_chunks && rc == 0 && !pageable; c++) {
        const int si = (int)(c % (u64)n_streams);
        const u64 a0 = c * chunk_arrays;
        const u64 cnt = (a0 + chunk_arrays <= batch) ? chunk_arrays : batch - a0;
        const u64 bytes = cnt * array_words * sizeof(u64);
        uint64_t *h = data + a0 * array_words;
        if (cudaMemcpyAsync(bufs[si], h, bytes, cudaMemcpyHostToDevice, streams[si]) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync H2D", __LINE__);
            break;
        }
        rc = tf21_ntt_dev(bufs[si], n, width, cnt, inverse, (tf21_stream_t)streams[si]);
        if (rc) break;
        if (cudaMemcpyAsync(h, bufs[si], bytes, cudaMemcpyDeviceToHost, streams[si]) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync D2H", __LINE__);
    }
    for (int i = 0; i < n_streams && rc == 0; i++)
        if (streams[i] && cudaStreamSynchronize(streams[i]) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "cudaStreamSynchronize", __LINE__);
    cleanup();
    return rc;
}

int tf21_ntt(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch) { return host_ntt(data, n, width, batch, 0); }
int tf21_intt(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch) { return host_ntt(data, n, width, batch, 1); }

// ---- coset -----------------------------------------------------------------------------------------
int tf21_coset_evaluate_dev(const uint64_t *d_coeffs, uint64_t n_coeffs, uint32_t width, uint64_t offset_raw,
                            uint64_t order, uint64_t *d_out, tf21_stream_t stream) {
    TF21_TRY(check_ntt_len(order, width));
    if (n_coeffs > order) return TF21_E_ORDER_LE_DEGREE;  // polynomial.rs:1388-1392
    if (order == 0) return 0;
    if (!d_out || (n_coeffs && !d_coeffs)) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    if (n_coeffs == 0) {
        TF21_CUDA(cudaMemsetAsync(d_out, 0, order * width * sizeof(u64), st));
        return 0;
    }
    // c_i * offset^i (Polynomial::scale, polynomial.rs:760-773) on load, zero padding to `order`
    u64 g = hgl_from_raw(offset_raw);
    ScaleTab pre;
    TF21_TRY(split_locked(*t, g, 1, n_coeffs, &pre));
    return ntt_run_locked(*t, d_coeffs, n_coeffs, d_out, order, width, 1, 0, pre, NO_SCALE, 0, st);
}

int tf21_coset_interpolate_dev(const uint64_t *d_values, uint64_t n, uint32_t width, uint64_t offset_raw,
                               uint64_t *d_coeffs_out, tf21_stream_t stream) {
    TF21_TRY(check_ntt_len(n, width));
    if (n == 0) return 0;
    if (!d_values || !d_coeffs_out) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    // u = iNTT(values); c_i = u_i * offset^-i  (polynomial.rs:1912-1917); n^-1 folded into the table
    u64 g_inv = hgl_inv(hgl_from_raw(offset_raw));  // offset.inverse() panics on zero in the reference
    ScaleTab post;
    TF21_TRY(split_locked(*t, g_inv, hgl_inv(n % GL_P), n, &post));
    if (n == 1) {
        TF21_CUDA(cudaMemcpyAsync(d_coeffs_out, d_values, width * sizeof(u64), cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)stream));
        return 0;
    }
    return ntt_run_locked(*t, d_values, n, d_coeffs_out, n, width, 1, 1, NO_SCALE, post, 0, (cudaStream_t)stream);
}

int tf21_coset_lde_dev(const uint64_t *d_values, uint64_t n_in, uint64_t offset_in_raw, uint64_t n_out,
                       uint64_t offset_out_raw, uint32_t width, uint64_t *d_out, tf21_stream_t stream) {
    TF21_TRY(check_ntt_len(n_in, width));
    TF21_TRY(check_ntt_len(n_out, width));
    if (n_in == 0 || n_out < n_in) return TF21_E_BAD_ARG;
    if (!d_values || !d_out) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    // coefficients c_i = iNTT(v)_i * n_in^-1 * (g_out / g_in)^i, then zero-extended NTT of size n_out
    u64 ratio = hgl_mul(hgl_from_raw(offset_out_raw), hgl_inv(hgl_from_raw(offset_in_raw)));
    Scratch coeffs(st);
    TF21_TRY(coeffs.alloc(n_in * width));
    if (n_in == 1) {
        TF21_CUDA(cudaMemcpyAsync(coeffs.p, d_values, width * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    } else {
        ScaleTab post;
        TF21_TRY(split_locked(*t, ratio, hgl_inv(n_in % GL_P), n_in, &post));
        TF21_TRY(ntt_run_locked(*t, d_values, n_in, coeffs.p, n_in, width, 1, 1, NO_SCALE, post, 0, st));
    }
    return ntt_run_locked(*t, coeffs.p, n_in, d_out, n_out, width, 1, 0, NO_SCALE, NO_SCALE, 0, st);
}

int tf21_coset_evaluate(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t width, uint64_t offset_raw,
                        uint64_t order, uint64_t *out) {
    TF21_TRY(check_ntt_len(order, width));
    // degree = index of the last non-zero coefficient (Polynomial::new strips trailing zeros)
    u64 live = n_coeffs;
    while (live > 0) {
        bool nz = false;
        for (u32 c = 0; c < width; c++) nz |= coeffs[(live - 1) * width + c] != 0;
        if (nz) break;
        live--;
    }
    if (!(order > live - 1 || live == 0)) return TF21_E_ORDER_LE_DEGREE;  // order > degree, :1388-1392
    if (order == 0) return 0;
    DevBuf in, outb;
    TF21_TRY(in.alloc(live * width));
    TF21_TRY(outb.alloc(order * width));
    if (live) TF21_TRY(copy_h2d(in.p, coeffs, live * width * sizeof(u64), nullptr));
    TF21_TRY(tf21_coset_evaluate_dev(in.p, live, width, offset_raw, order, outb.p, nullptr));
    TF21_TRY(copy_d2h(out, outb.p, order * width * sizeof(u64), nullptr));
    return 0;
}

int tf21_coset_interpolate(const uint64_t *values, uint64_t n, uint32_t width, uint64_t offset_raw,
                           uint64_t *coeffs_out) {
    TF21_TRY(check_ntt_len(n, width));
    if (n == 0) return 0;
    DevBuf in, outb;
    TF21_TRY(in.alloc(n * width));
    TF21_TRY(outb.alloc(n * width));
    TF21_TRY(copy_h2d(in.p, values, n * width * sizeof(u64), nullptr));
    TF21_TRY(tf21_coset_interpolate_dev(in.p, n, width, offset_raw, outb.p, nullptr));
    TF21_TRY(copy_d2h(coeffs_out, outb.p, n * width * sizeof(u64), nullptr));
    return 0;
}

int tf21_coset_lde(const uint64_t *values, uint64_t n_in, uint64_t offset_in_raw, uint64_t n_out,
                   uint64_t offset_out_raw, uint32_t width, uint64_t *out) {
    TF21_TRY(check_ntt_len(n_in, width));
    TF21_TRY(check_ntt_len(n_out, width));
    if (n_in == 0 || n_out < n_in) return TF21_E_BAD_ARG;
    DevBuf in, outb;
    TF21_TRY(in.alloc(n_in * width));
    TF21_TRY(outb.alloc(n_out * width));
    TF21_TRY(copy_h2d(in.p, values, n_in * width * sizeof(u64), nullptr));
    TF21_TRY(tf21_coset_lde_dev(in.p, n_in, offset_in_raw, n_out, offset_out_raw, width, outb.p, nullptr));
    TF21_TRY(copy_d2h(out, outb.p, n_out * width * sizeof(u64), nullptr));
    return 0;
}

// ---- polynomial multiplication (next wave) -----------------------------------------
int tf21_poly_mul_dev(const uint64_t *d_a, uint64_t n_a, const uint64_t *d_b, uint64_t n_b, uint32_t width,
                      uint64_t *d_out, tf21_stream_t stream) {
    if (width != 1 && width != 3) return TF21_E_BAD_ARG;
    if (n_a == 0 || n_b == 0) return 0;  // zero polynomial: nothing to write (polynomial.rs:909-911)
    if (!d_a || !d_b || !d_out) return TF21_E_BAD_ARG;
    const u64 len = n_a + n_b - 1;
    u64 order = 1;
    while (order < len) order <<= 1;  // (degree + 1).next_power_of_two(), polynomial.rs:912
    TF21_TRY(check_ntt_len(order, width));
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    const bool square = d_a == d_b && n_a == n_b;  // fast_square (polynomial.rs:780-802): one forward transform
    Scratch l(st), r(st);
    TF21_TRY(l.alloc(order * width));
    if (!square || order != len) TF21_TRY(r.alloc(order * width));
    // resize(order, ZERO) + ntt, fused: the transforms read only the n_a / n_b coefficients that exist
    TF21_TRY(ntt_run_locked(*t, d_a, n_a, l.p, order, width, 1, 0, NO_SCALE, NO_SCALE, 0, st));
    if (!square) TF21_TRY(ntt_run_locked(*t, d_b, n_b, r.p, order, width, 1, 0, NO_SCALE, NO_SCALE, 0, st));
    const u64 rinv = hgl_inv(GL_EPS);  // 2^-64 mod p
    TF21_LAUNCH(hadamard_kernel, grid_for(order, 256), 256, 0, st, l.p, square ? l.p : r.p, order, width, rinv);
    const u64 post_scalar = hgl_inv(order % GL_P);
    if (order == len) {
        TF21_TRY(ntt_run_locked(*t, l.p, order, d_out, order, width, 1, 1, NO_SCALE, NO_SCALE,
                                order > 1 ? post_scalar : 0, st));
        if (order == 1) {  // identity transform: still canonicalise the lazy product
            TF21_TRY(tf21_selftest_field_dev(3, d_out, d_out, d_out, width, stream));
        }
    } else {
        TF21_TRY(ntt_run_locked(*t, l.p, order, r.p, order, width, 1, 1, NO_SCALE, NO_SCALE, post_scalar, st));
        TF21_CUDA(cudaMemcpyAsync(d_out, r.p, len * width * sizeof(u64), cudaMemcpyDeviceToDevice, st));  // truncate
    }
    return 0;
}

int tf21_poly_square_dev(const uint64_t *d_a, uint64_t n_a, uint32_t width, uint64_t *d_out, tf21_stream_t stream) {
    return tf21_poly_mul_dev(d_a, n_a, d_a, n_a, width, d_out, stream);
}

int tf21_poly_square(const uint64_t *a, uint64_t n_a, uint32_t width, uint64_t *out) {
    if (width != 1 && width != 3) return TF21_E_BAD_ARG;
    if (n_a == 0) return 0;
    if (!a || !out) return TF21_E_BAD_ARG;
    const u64 len = 2 * n_a - 1;
    DevBuf da, dout;
    TF21_TRY(da.alloc(n_a * width));
    TF21_TRY(dout.alloc(len * width));
    TF21_TRY(copy_h2d(da.p, a, n_a * width * sizeof(u64), nullptr));
    TF21_TRY(tf21_poly_square_dev(da.p, n_a, width, dout.p, nullptr));
    TF21_TRY(copy_d2h(out, dout.p, len * width * sizeof(u64), nullptr));
    return 0;
}

int tf21_poly_mul(const uint64_t *a, uint64_t n_a, const uint64_t *b, uint64_t n_b, uint32_t width, uint64_t *out) {
    if (width != 1 && width != 3) return TF21_E_BAD_ARG;
    if (n_a == 0 || n_b == 0) return 0;
    if (!a || !b || !out) return TF21_E_BAD_ARG;
    const u64 len = n_a + n_b - 1;
    DevBuf da, db, dout;
    TF21_TRY(da.alloc(n_a * width));
    TF21_TRY(db.alloc(n_b * width));
    TF21_TRY(dout.alloc(len * width));
    TF21_TRY(copy_h2d(da.p, a, n_a * width * sizeof(u64), nullptr));
    TF21_TRY(copy_h2d(db.p, b, n_b * width * sizeof(u64), nullptr));
    TF21_TRY(tf21_poly_mul_dev(da.p, n_a, db.p, n_b, width, dout.p, nullptr));
    TF21_TRY(copy_d2h(out, dout.p, len * width * sizeof(u64), nullptr));
    return 0;
}

// ---- Polynomial::reduce_by_ntt_friendly_modulus (next wave; polynomial.rs:1087-1148) --------
// f mod (X^(chunk + tail) + shift(X)), deg shift < tail, with shift given in the NTT domain (domain_length =
// chunk + tail values).  The chunk loop of the reference stays (each step needs the window of the previous one),
// every step is device work: zero-extended NTT of the top of the window, Hadamard product with shift_ntt,
// inverse NTT, window update.
int tf21_poly_reduce_by_ntt_friendly_modulus(const uint64_t *coeffs, uint64_t n_coeffs, uint32_t width,
                                             const uint64_t *shift_ntt, uint64_t domain_length,
                                             uint64_t tail_length, uint64_t *out, uint64_t *n_out) {
    if (width != 1 && width != 3) return TF21_E_BAD_ARG;
    if (domain_length == 0 || (domain_length & (domain_length - 1))) return TF21_E_LEN_NOT_POW2;  // :1094 assert
    TF21_TRY(check_ntt_len(domain_length, width));
    if (!n_out || tail_length >= domain_length || !shift_ntt || (n_coeffs && (!coeffs || !out))) return TF21_E_BAD_ARG;
    const u64 chunk = domain_length - tail_length, w = width;
    if (n_coeffs < domain_length) {  // :1097-1099: nothing to reduce
        if (n_coeffs) std::memcpy(out, coeffs, n_coeffs * w * sizeof(u64));
        *n_out = n_coeffs;
        return 0;
    }
    *n_out = domain_length;
    const u64 n_chunks = (n_coeffs - domain_length + chunk - 1) / chunk;  // :1100-1101
    const u64 range_start = n_chunks * chunk;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = nullptr;
    DevBuf dc, dshift, win[2], prod;
    TF21_TRY(dc.alloc(n_coeffs * w));
    TF21_TRY(dshift.alloc(domain_length * w));
    TF21_TRY(win[0].alloc(domain_length * w));
    TF21_TRY(win[1].alloc(domain_length * w));
    TF21_TRY(prod.alloc(domain_length * w));
    TF21_TRY(copy_h2d(dc.p, coeffs, n_coeffs * w * sizeof(u64), nullptr));
    TF21_TRY(copy_h2d(dshift.p, shift_ntt, domain_length * w * sizeof(u64), nullptr));
    // working_window = coefficients[range_start..] zero-padded to chunk + tail (:1103-1109)
    TF21_CUDA(cudaMemsetAsync(win[0].p, 0, domain_length * w * sizeof(u64), st));
    if (range_start < n_coeffs)
        TF21_CUDA(cudaMemcpyAsync(win[0].p, dc.p + range_start * w, (n_coeffs - range_start) * w * sizeof(u64),
                                  cudaMemcpyDeviceToDevice, st));
    const u64 rinv = hgl_inv(GL_EPS);
    const u64 post_scalar = hgl_inv(domain_length % GL_P);
    int cur = 0;
    for (u64 ci = n_chunks; ci-- > 0;) {
        // product = iNTT( NTT( window[tail..] | 0 x tail ) .* shift_ntt )   (:1112-1124)
        if (domain_length > 1) {
            TF21_TRY(ntt_run_locked(*t, win[cur].p + tail_length * w, chunk, prod.p, domain_length, width, 1, 0, NO_SCALE,
                                    NO_SCALE, 0, st));
        } else {
            TF21_CUDA(cudaMemcpyAsync(prod.p, win[cur].p + tail_length * w, w * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        }
        TF21_LAUNCH(hadamard_kernel, grid_for(domain_length, 256), 256, 0, st, prod.p, dshift.p, domain_length, width, rinv);
        if (domain_length > 1)
            TF21_TRY(ntt_run_locked(*t, prod.p, domain_length, prod.p, domain_length, width, 1, 1, NO_SCALE, NO_SCALE,
                                    post_scalar, st));
        // window = [coefficients of chunk ci | old window[0..tail)] - product   (:1126-1143)
        TF21_LAUNCH(reduce_window_kernel, grid_for(domain_length * w, 256), 256, 0, st, dc.p + ci * chunk * w, win[cur].p,
                    prod.p, chunk * w, domain_length * w, win[cur ^ 1].p);
        cur ^= 1;
    }
    TF21_TRY(copy_d2h(out, win[cur].p, domain_length * w * sizeof(u64), nullptr));
    return 0;
}

// ---- Polynomial::clean_divide (next wave; polynomial.rs:2358-2413) ---------------------------
// q = a / b for BFieldElement polynomials whose division leaves no remainder: both are evaluated on a coset of
// order next_power_of_two(deg a + 1), divided point-wise and interpolated back.  The reference moves to the
// extension field to dodge roots of the divisor on the coset; here the coset offset stays in the base field and
// is changed when a root is hit (the quotient does not depend on it).
int tf21_poly_clean_divide(const uint64_t *a, uint64_t n_a, const uint64_t *b, uint64_t n_b, uint64_t *q_out,
                           uint64_t *n_q) {
    if (!n_q || (n_a && !a) || (n_b && !b)) return TF21_E_BAD_ARG;
    while (n_a && a[n_a - 1] == 0) n_a--;  // degree() ignores trailing zeros (polynomial.rs:181-191)
    while (n_b && b[n_b - 1] == 0) n_b--;
    if (n_b == 0) return TF21_E_DIVISION_BY_ZERO;  // "divisor should be non-zero"
    if (n_a < n_b) {  // clean division of a lower-degree dividend: only the zero polynomial qualifies
        *n_q = 0;
        return 0;
    }
    const u64 len = n_a - n_b + 1;
    *n_q = len;
    if (!q_out) return TF21_E_BAD_ARG;
    u64 order = 1;
    while (order < n_a) order <<= 1;
    TF21_TRY(check_ntt_len(order, 1));
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    DevBuf da, db, ea, eb, flag;
    TF21_TRY(da.alloc(n_a));
    TF21_TRY(db.alloc(n_b));
    TF21_TRY(ea.alloc(order));
    TF21_TRY(eb.alloc(order));
    TF21_TRY(flag.alloc(1));
    TF21_TRY(copy_h2d(da.p, a, n_a * sizeof(u64), nullptr));
    TF21_TRY(copy_h2d(db.p, b, n_b * sizeof(u64), nullptr));
    u64 offset = 7;  // BFieldElement::generator()
    for (int attempt = 0; attempt < 8; attempt++, offset = hgl_mul(offset, 7)) {
        const u64 offset_raw = hgl_to_raw(offset);
        TF21_TRY(tf21_coset_evaluate_dev(da.p, n_a, 1, offset_raw, order, ea.p, nullptr));
        TF21_TRY(tf21_coset_evaluate_dev(db.p, n_b, 1, offset_raw, order, eb.p, nullptr));
        TF21_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(u64), nullptr));
        TF21_LAUNCH(pointwise_divide_kernel, grid_for(order, 256), 256, 0, (cudaStream_t) nullptr, ea.p, eb.p, order,
                    (u32 *)flag.p);
        u64 hit = 0;
        TF21_TRY(copy_d2h(&hit, flag.p, sizeof(u64), nullptr));
        if (hit) continue;  // the coset contains a root of the divisor: next offset
        TF21_TRY(tf21_coset_interpolate_dev(ea.p, order, 1, offset_raw, eb.p, nullptr));
        TF21_TRY(copy_d2h(q_out, eb.p, len * sizeof(u64), nullptr));
        return 0;
    }
    return TF21_E_BAD_ARG;
}

// ---- out-of-domain evaluation / coset extrapolation (next wave) ---------------------------
int tf21_poly_evaluate_batch_dev(const uint64_t *d_polys, uint64_t n, uint64_t n_polys, uint32_t width,
                                 const uint64_t *points, uint64_t n_points, uint64_t *d_out, tf21_stream_t stream) {
    if (width != 1 && width != 3) return TF21_E_BAD_ARG;
    if (n_polys == 0 || n_points == 0) return 0;
    if ((n && !d_polys) || !points || !d_out) return TF21_E_BAD_ARG;
    if (n_points > 0x7fffffffull || n_polys > 65535) return TF21_E_LEN_TOO_LARGE;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<u64> vals(n_points * width);
    for (u64 i = 0; i < vals.size(); i++) vals[i] = hgl_from_raw(points[i]);  // canonical values of the points
    Scratch dpts(st);
    TF21_TRY(dpts.alloc(vals.size()));
    TF21_CUDA(cudaMemcpyAsync(dpts.p, vals.data(), vals.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    dim3 grid((unsigned)n_points, (unsigned)n_polys);
    if (width == 1)
        TF21_LAUNCH_NAMED("poly_eval_kernel", poly_eval_kernel<1>, grid, kEvalThreads, 0, st, d_polys, n, dpts.p, (u32)n_points, d_out);
    else
        TF21_LAUNCH_NAMED("poly_eval_kernel", poly_eval_kernel<3>, grid, kEvalThreads, 0, st, d_polys, n, dpts.p, (u32)n_points, d_out);
    return 0;
}

// Polynomial::par_batch_coset_extrapolate (polynomial.rs:2255-2331): every codeword is interpolated on the coset
// (batched iNTT with offset^-i n^-1 fused into its last pass) and evaluated in every point.  The reference picks
// between two algorithms by the number of points; both compute the values of the unique interpolant.
int tf21_batch_coset_extrapolate_dev(uint64_t offset_raw, uint64_t codeword_length, const uint64_t *d_codewords,
                                     uint64_t n_codewords, uint32_t width, const uint64_t *points, uint64_t n_points,
                                     uint64_t *d_out, tf21_stream_t stream) {
    TF21_TRY(check_ntt_len(codeword_length, width));
    if (codeword_length == 0) return TF21_E_BAD_ARG;
    if (n_codewords == 0 || n_points == 0) return 0;
    if (!d_codewords || !points || !d_out) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    const u64 n = codeword_length;
    Scratch coeffs(st);
    TF21_TRY(coeffs.alloc(n * width * n_codewords));
    if (n == 1) {
        TF21_CUDA(cudaMemcpyAsync(coeffs.p, d_codewords, n_codewords * width * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    } else {
        u64 g_inv = hgl_inv(hgl_from_raw(offset_raw));
        ScaleTab post;
        TF21_TRY(split_locked(*t, g_inv, hgl_inv(n % GL_P), n, &post));
        TF21_TRY(ntt_run_locked(*t, d_codewords, n, coeffs.p, n, width, n_codewords, 1, NO_SCALE, post, 0, st));
    }
    // y-dimension of the grid is limited to 65535 polynomials per launch
    for (u64 first = 0; first < n_codewords; first += 65535) {
        const u64 cnt = n_codewords - first < 65535 ? n_codewords - first : 65535;
        TF21_TRY(tf21_poly_evaluate_batch_dev(coeffs.p + first * n * width, n, cnt, width, points, n_points,
                                              d_out + first * n_points * width, stream));
    }
    return 0;
}

int tf21_batch_coset_extrapolate(uint64_t offset_raw, uint64_t codeword_length, const uint64_t *codewords,
                                 uint64_t n_codewords, uint32_t width, const uint64_t *points, uint64_t n_points,
                                 uint64_t *out) {
    TF21_TRY(check_ntt_len(codeword_length, width));
    if (codeword_length == 0) return TF21_E_BAD_ARG;
    if (n_codewords == 0 || n_points == 0) return 0;
    if (!codewords || !points || !out) return TF21_E_BAD_ARG;
    DevBuf dc, dout;
    const u64 words = codeword_length * width * n_codewords;
    TF21_TRY(dc.alloc(words));
    TF21_TRY(dout.alloc(n_codewords * n_points * width));
    TF21_TRY(copy_h2d(dc.p, codewords, words * sizeof(u64), nullptr));
    TF21_TRY(tf21_batch_coset_extrapolate_dev(offset_raw, codeword_length, dc.p, n_codewords, width, points, n_points,
                                              dout.p, nullptr));
    TF21_TRY(copy_d2h(out, dout.p, n_codewords * n_points * width * sizeof(u64), nullptr));
    return 0;
}

// ---- Tip5 ------------------------------------------------------------------------------------------
int tf21_tip5_permute_dev(uint64_t *d_states, uint64_t count, tf21_stream_t stream) {
    if (count && (!d_states || misaligned16(d_states))) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    return launch_permute(d_states, count, (cudaStream_t)stream);
}
int tf21_tip5_hash_10_dev(const uint64_t *d_in, uint64_t count, uint64_t *d_out, tf21_stream_t stream) {
    if (count && (!d_in || !d_out || misaligned16(d_in))) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    return launch_hash10(d_in, count, d_out, (cudaStream_t)stream);
}
int tf21_tip5_hash_rows_dev(const uint64_t *d_rows, uint64_t row_len, uint64_t n_rows, uint64_t *d_out,
                            tf21_stream_t stream) {
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    return launch_hash_rows(d_rows, row_len, n_rows, row_len, 1, d_out, (cudaStream_t)stream);
}
int tf21_tip5_hash_columns_dev(const uint64_t *d_cols, uint64_t n_rows, uint64_t n_cols, uint64_t col_stride_words,
                               uint64_t *d_out, tf21_stream_t stream) {
    if (n_rows == 0) return 0;
    if (!d_out || (n_cols && !d_cols)) return TF21_E_BAD_ARG;
    if (n_cols > 1 && col_stride_words < n_rows) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    return launch_hash_rows(d_cols, n_cols, n_rows, 1, col_stride_words, d_out, (cudaStream_t)stream);
}

// Tip5::sample_indices (tip5/mod.rs:636-656) on a sponge state held by the caller (16 raw words, updated)
int tf21_tip5_sample_indices(uint64_t *state, uint32_t upper_bound, uint64_t num_indices, uint32_t *out) {
    if (upper_bound == 0 || (upper_bound & (upper_bound - 1))) return TF21_E_LEN_NOT_POW2;  // assert!(is_power_of_two)
    if (!state || (num_indices && !out)) return TF21_E_BAD_ARG;
    if (num_indices == 0) return 0;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    DevBuf ds, dout;
    TF21_TRY(ds.alloc(16));
    TF21_TRY(dout.alloc((num_indices + 1) / 2));
    TF21_TRY(copy_h2d(ds.p, state, 16 * sizeof(u64), nullptr));
    TF21_LAUNCH(tip5_sample_indices_kernel, 1, 32, 0, (cudaStream_t) nullptr, ds.p, upper_bound, num_indices,
                hgl_inv(GL_EPS), (u32 *)dout.p);
    TF21_TRY(copy_d2h(out, dout.p, num_indices * sizeof(u32), nullptr));
    TF21_TRY(copy_d2h(state, ds.p, 16 * sizeof(u64), nullptr));
    return 0;
}

int tf21_tip5_permute(uint64_t *states, uint64_t count) {
    if (count == 0) return 0;
    if (!states) return TF21_E_BAD_ARG;
    DevBuf b;
    TF21_TRY(b.alloc(16 * count));
    TF21_TRY(copy_h2d(b.p, states, 16 * count * sizeof(u64), nullptr));
    TF21_TRY(tf21_tip5_permute_dev(b.p, count, nullptr));
    TF21_TRY(copy_d2h(states, b.p, 16 * count * sizeof(u64), nullptr));
    return 0;
}

int tf21_tip5_hash_10(const uint64_t *in, uint64_t count, uint64_t *out) {
    if (count == 0) return 0;
    if (!in || !out) return TF21_E_BAD_ARG;
    DevBuf bi, bo;
    TF21_TRY(bi.alloc(10 * count));
    TF21_TRY(bo.alloc(5 * count));
    TF21_TRY(copy_h2d(bi.p, in, 10 * count * sizeof(u64), nullptr));
    TF21_TRY(tf21_tip5_hash_10_dev(bi.p, count, bo.p, nullptr));
    TF21_TRY(copy_d2h(out, bo.p, 5 * count * sizeof(u64), nullptr));
    return 0;
}

int tf21_tip5_hash_pairs(const uint64_t *pairs, uint64_t count, uint64_t *out) {
    return tf21_tip5_hash_10(pairs, count, out);  // hash_pair is hash_10 of left | right, tip5/mod.rs:577-586
}

int tf21_tip5_hash_rows(const uint64_t *rows, uint64_t row_len, uint64_t n_rows, uint64_t *out) {
    if (n_rows == 0) return 0;
    if (!out || (row_len && !rows)) return TF21_E_BAD_ARG;
    DevBuf bi, bo;
    TF21_TRY(bi.alloc(row_len * n_rows));
    TF21_TRY(bo.alloc(5 * n_rows));
    if (row_len) TF21_TRY(copy_h2d(bi.p, rows, row_len * n_rows * sizeof(u64), nullptr));
    TF21_TRY(tf21_tip5_hash_rows_dev(bi.p, row_len, n_rows, bo.p, nullptr));
    TF21_TRY(copy_d2h(out, bo.p, 5 * n_rows * sizeof(u64), nullptr));
    return 0;
}

int tf21_tip5_hash_varlen(const uint64_t *in, uint64_t len, uint64_t out[5]) {
    return tf21_tip5_hash_rows(in, len, 1, out);
}

// ---- Merkle ----------------------------------------------------------------------------------------
int tf21_merkle_build_dev(const uint64_t *d_leafs, uint64_t n_leafs, uint64_t *d_nodes_out, tf21_stream_t stream) {
    TF21_TRY(check_leaf_count(n_leafs));
    if (!d_leafs || !d_nodes_out) return TF21_E_BAD_ARG;
    if (misaligned16(d_leafs) || misaligned16(d_nodes_out)) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    return merkle_build_dev(d_leafs, n_leafs, d_nodes_out, (cudaStream_t)stream);
}

int tf21_merkle_root_dev(const uint64_t *d_leafs, uint64_t n_leafs, uint64_t *d_root_out, tf21_stream_t stream) {
    TF21_TRY(check_leaf_count(n_leafs));
    if (!d_leafs || !d_root_out || misaligned16(d_leafs)) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    if (n_leafs == 1) {
        TF21_CUDA(cudaMemcpyAsync(d_root_out, d_leafs, 5 * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    // heap-indexed upper half only: upper[cnt..2cnt) for cnt <= n/2; the leaves are read in place
    Scratch upper(st);
    TF21_TRY(upper.alloc(5 * n_leafs));
    u64 cnt = n_leafs / 2;
    TF21_TRY(launch_hash10(d_leafs, cnt, upper.p + 5 * cnt, st));
    if (cnt > 1) TF21_TRY(launch_merkle_levels(upper.p, cnt, st));
    TF21_CUDA(cudaMemcpyAsync(d_root_out, upper.p + 5, 5 * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int tf21_merkle_scatter_subtree_dev(const uint64_t *d_local_nodes, uint64_t n_local_leafs, uint64_t shard,
                                    uint64_t n_shards, uint64_t *d_global_nodes, tf21_stream_t stream) {
    TF21_TRY(check_leaf_count(n_local_leafs));
    TF21_TRY(check_leaf_count(n_shards));
    if (shard >= n_shards || !d_local_nodes || !d_global_nodes) return TF21_E_BAD_ARG;
    u64 total = (2 * n_local_leafs - 1) * 5;
    TF21_LAUNCH(merkle_scatter_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, d_local_nodes,
                n_local_leafs, shard, n_shards, d_global_nodes);
    return 0;
}

// Merkle root of d_leafs[0..n) into d_root_out using caller-provided scratch of 5*n words (n a power of two)
static int merkle_root_with_scratch(const u64 *d_leafs, u64 n, u64 *d_root_out, u64 *scratch, cudaStream_t st) {
    if (n == 1) {
        TF21_CUDA(cudaMemcpyAsync(d_root_out, d_leafs, 5 * sizeof(u64), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    u64 cnt = n / 2;
    TF21_TRY(launch_hash10(d_leafs, cnt, scratch + 5 * cnt, st));
    if (cnt > 1) TF21_TRY(launch_merkle_levels(scratch, cnt, st));
    TF21_CUDA(cudaMemcpyAsync(d_root_out, scratch + 5, 5 * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---- next wave: authentication structures ------------------------------------------
// merkle_tree.rs:449-504.  Host-side set logic: `needed` = siblings along every leaf-to-root path,
// `computable` = the path nodes themselves; result = needed \ computable, descending.
int tf21_merkle_auth_structure_node_indices(uint64_t n_leafs, const uint64_t *leaf_indices, uint64_t n_indices,
                                            uint64_t *out, uint64_t capacity, uint64_t *count) {
    if (n_leafs == 0 || (n_leafs & (n_leafs - 1))) return TF21_E_INCORRECT_NUMBER_OF_LEAFS;  // :467-469
    if (!count || (n_indices && !leaf_indices)) return TF21_E_BAD_ARG;
    std::vector<u64> needed, computable;
    for (u64 k = 0; k < n_indices; k++) {
        if (leaf_indices[k] >= n_leafs) return TF21_E_LEAF_INDEX_INVALID;  // :487-489
        for (u64 node = leaf_indices[k] + n_leafs; node > 1; node >>= 1) {
            computable.push_back(node);
            needed.push_back(node ^ 1);
        }
    }
    std::sort(needed.begin(), needed.end());
    needed.erase(std::unique(needed.begin(), needed.end()), needed.end());
    std::sort(computable.begin(), computable.end());
    u64 c = 0;
    for (auto it = needed.rbegin(); it != needed.rend(); ++it) {
        if (std::binary_search(computable.begin(), computable.end(), *it)) continue;
        if (c < capacity && out) out[c] = *it;
        c++;
    }
    *count = c;
    return (c > capacity || (c && !out)) ? TF21_E_CAPACITY : 0;
}

int tf21_merkle_authentication_structure_dev(const uint64_t *d_nodes, uint64_t n_leafs, const uint64_t *leaf_indices,
                                             uint64_t n_indices, uint64_t *d_out, uint64_t capacity,
                                             uint64_t *count, tf21_stream_t stream) {
    if (!d_nodes || !count) return TF21_E_BAD_ARG;
    u64 c = 0;
    int rc = tf21_merkle_auth_structure_node_indices(n_leafs, leaf_indices, n_indices, nullptr, 0, &c);
    if (rc != 0 && rc != TF21_E_CAPACITY) return rc;
    *count = c;
    if (c > capacity || (c && !d_out)) return TF21_E_CAPACITY;
    if (c == 0) return 0;
    std::vector<u64> idx(c);
    TF21_TRY(tf21_merkle_auth_structure_node_indices(n_leafs, leaf_indices, n_indices, idx.data(), c, &c));
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    Scratch didx(st);
    TF21_TRY(didx.alloc(c));
    // pageable source: the copy is staged before cudaMemcpyAsync returns, so `idx` may go out of scope
    TF21_CUDA(cudaMemcpyAsync(didx.p, idx.data(), c * sizeof(u64), cudaMemcpyHostToDevice, st));
    TF21_LAUNCH(merkle_gather_kernel, grid_for(5 * c, 256), 256, 0, st, d_nodes, didx.p, c, d_out);
    return 0;
}

// MerkleTree::{sequential,par}_authentication_structure_from_leafs (merkle_tree.rs:514-542): the reference
// computes one frugal root per needed node; here the tree is built once on the device and gathered from.
int tf21_merkle_authentication_structure_from_leafs(const uint64_t *leafs, uint64_t n_leafs,
                                                    const uint64_t *leaf_indices, uint64_t n_indices,
                                                    uint64_t *out, uint64_t capacity, uint64_t *count) {
    if (!count) return TF21_E_BAD_ARG;
    u64 c = 0;
    int rc = tf21_merkle_auth_structure_node_indices(n_leafs, leaf_indices, n_indices, nullptr, 0, &c);
    if (rc != 0 && rc != TF21_E_CAPACITY) return rc;
    *count = c;
    if (c > capacity || (c && !out)) return TF21_E_CAPACITY;
    if (c == 0) return 0;
    if (!leafs) return TF21_E_BAD_ARG;
    DevBuf bl, bn, bo;
    TF21_TRY(bl.alloc(5 * n_leafs));
    TF21_TRY(bn.alloc(10 * n_leafs));
    TF21_TRY(bo.alloc(5 * c));
    TF21_TRY(copy_h2d(bl.p, leafs, 5 * n_leafs * sizeof(u64), nullptr));
    TF21_TRY(tf21_merkle_build_dev(bl.p, n_leafs, bn.p, nullptr));
    TF21_TRY(tf21_merkle_authentication_structure_dev(bn.p, n_leafs, leaf_indices, n_indices, bo.p, c, &c, nullptr));
    TF21_TRY(copy_d2h(out, bo.p, 5 * c * sizeof(u64), nullptr));
    return 0;
}

// ---- next wave: MMR bulk operations -------------------------------------------------
// MmrAccumulator::peaks_from_leafs (mmr/mmr_accumulator.rs:96-115): the peaks are the Merkle roots of the
// power-of-two runs of the binary expansion of n_leafs, most significant first; an odd count leaves the
// last leaf as its own peak.  Each run is one level-batched build.
int tf21_mmr_peaks_from_leafs_dev(const uint64_t *d_leafs, uint64_t n_leafs, uint64_t *d_peaks_out,
                                  uint64_t *n_peaks, tf21_stream_t stream) {
    if (!n_peaks) return TF21_E_BAD_ARG;
    *n_peaks = (u64)__builtin_popcountll(n_leafs);
    if (n_leafs == 0) return 0;
    if (!d_leafs || !d_peaks_out || misaligned16(d_leafs)) return TF21_E_BAD_ARG;
    if (n_leafs > (1ull << 40)) return TF21_E_ALLOC;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned top = ilog2_u64(n_leafs);
    Scratch scratch(st);
    if (top >= 1) TF21_TRY(scratch.alloc(5ull << top));
    u64 off = 0, k = 0;
    for (int b = (int)top; b >= 0; b--) {
        if (!((n_leafs >> b) & 1)) continue;
        TF21_TRY(merkle_root_with_scratch(d_leafs + 5 * off, 1ull << b, d_peaks_out + 5 * k, scratch.p, st));
        off += 1ull << b;
        k++;
    }
    return 0;
}

int tf21_mmr_peaks_from_leafs(const uint64_t *leafs, uint64_t n_leafs, uint64_t *peaks_out, uint64_t *n_peaks) {
    if (!n_peaks) return TF21_E_BAD_ARG;
    *n_peaks = (u64)__builtin_popcountll(n_leafs);
    if (n_leafs == 0) return 0;
    if (!leafs || !peaks_out) return TF21_E_BAD_ARG;
    DevBuf bl, bp;
    TF21_TRY(bl.alloc(5 * n_leafs));
    TF21_TRY(bp.alloc(5 * 64));
    TF21_TRY(copy_h2d(bl.p, leafs, 5 * n_leafs * sizeof(u64), nullptr));
    TF21_TRY(tf21_mmr_peaks_from_leafs_dev(bl.p, n_leafs, bp.p, n_peaks, nullptr));
    TF21_TRY(copy_d2h(peaks_out, bp.p, 5 * *n_peaks * sizeof(u64), nullptr));
    return 0;
}

int tf21_mmr_bag_peaks_dev(const uint64_t *d_peaks, uint64_t n_peaks, uint64_t leaf_count, uint64_t *d_out,
                           tf21_stream_t stream) {
    if (!d_out || (n_peaks && !d_peaks) || n_peaks > 64) return TF21_E_BAD_ARG;
    DeviceTables *t;
    TF21_TRY(get_tables(&t));
    // BFieldCodec of u64: two 32-bit limbs, low first (bfield_codec.rs:122-128)
    const u64 lo = hgl_to_raw(leaf_count & 0xffffffffull), hi = hgl_to_raw(leaf_count >> 32);
    TF21_LAUNCH(mmr_bag_peaks_kernel, 1, 32, 0, (cudaStream_t)stream, d_peaks, (u32)n_peaks, lo, hi, d_out);
    return 0;
}

int tf21_mmr_bag_peaks(const uint64_t *peaks, uint64_t n_peaks, uint64_t leaf_count, uint64_t out[5]) {
    if (!out || (n_peaks && !peaks) || n_peaks > 64) return TF21_E_BAD_ARG;
    DevBuf bp, bo;
    TF21_TRY(bp.alloc(5 * n_peaks));
    TF21_TRY(bo.alloc(5));
    if (n_peaks) TF21_TRY(copy_h2d(bp.p, peaks, 5 * n_peaks * sizeof(u64), nullptr));
    TF21_TRY(tf21_mmr_bag_peaks_dev(bp.p, n_peaks, leaf_count, bo.p, nullptr));
    TF21_TRY(copy_d2h(out, bo.p, 5 * sizeof(u64), nullptr));
    return 0;
}

// ---- single-process sharding over the visible devices -------------------------------
// Columns and Merkle subtrees are independent (the reference's rayon split, ntt.rs:250-269 /
// merkle_tree.rs:165-212, 247-275), so one host thread per shard drives its own device; there is no data-path
// exchange: the only cross-shard step is finishing the top log2(n_shards) levels from the shard roots.
// Shard s runs on device s % n_devices, so the index algebra is testable on a single GPU.
static int visible_devices(int *n) {
    TF21_CUDA(cudaGetDeviceCount(n));
    *n = tf21_device_count();  // honours tf21_init(n_devices)
    return *n > 0 ? 0 : TF21_E_CUDA;
}

int tf21_ntt_sharded(uint64_t *data, uint64_t n, uint32_t width, uint64_t batch, int inverse, uint32_t n_shards) {
    TF21_TRY(check_ntt_len(n, width));
    if (n <= 1 || batch == 0) return 0;
    if (!data || n_shards == 0) return TF21_E_BAD_ARG;
    int n_dev = 0, prev = 0;
    TF21_TRY(visible_devices(&n_dev));
    cudaGetDevice(&prev);
    if (n_shards > batch) n_shards = (uint32_t)batch;
    std::vector<int> rc(n_shards, 0);
    std::vector<std::thread> workers;
    const u64 per = batch / n_shards, extra = batch % n_shards;
    u64 first = 0;
    for (uint32_t sh = 0; sh < n_shards; sh++) {
        const u64 cnt = per + (sh < extra ? 1 : 0);
        u64 *slice = data + first * n * width;
        first += cnt;
        workers.emplace_back([=, &rc] {
            if (cudaSetDevice((int)(sh % n_dev)) != cudaSuccess) {
                rc[sh] = TF21_E_CUDA;
                return;
            }
            rc[sh] = inverse ? tf21_intt(slice, n, width, cnt) : tf21_ntt(slice, n, width, cnt);
        });
    }
    for (auto &w : workers) w.join();
    cudaSetDevice(prev);
    for (int r : rc)
        if (r) return r;
    return 0;
}

// host path of the sharded build (no NCCL, or several shards per device): the shard trees come back to host
// memory, the cap is finished from the shard roots there
static int merkle_build_sharded_host(const uint64_t *leafs, uint64_t n_leafs, uint64_t *nodes_out, uint32_t n_shards,
                                     int n_dev) {
    int prev = 0;
    cudaGetDevice(&prev);
    const u64 local = n_leafs / n_shards;
    std::vector<int> rc(n_shards, 0);
    std::vector<std::vector<u64>> trees(n_shards);
    std::vector<std::thread> workers;
    for (uint32_t sh = 0; sh < n_shards; sh++) {
        workers.emplace_back([=, &rc, &trees] {
            if (cudaSetDevice((int)(sh % n_dev)) != cudaSuccess) {
                rc[sh] = TF21_E_CUDA;
                return;
            }
            trees[sh].resize(10 * local);
            rc[sh] = tf21_merkle_build(leafs + 5 * sh * local, local, trees[sh].data());
        });
    }
    for (auto &w : workers) w.join();
    cudaSetDevice(prev);
    for (int r : rc)
        if (r) return r;
    // local node j of shard sh (level width wd = 2^floor(log2 j)) sits at global n_shards * wd + sh * wd + (j - wd)
    for (uint32_t sh = 0; sh < n_shards; sh++)
        for (u64 wd = 1; wd <= local; wd <<= 1)
            std::memcpy(nodes_out + 5 * (n_shards * wd + sh * wd), trees[sh].data() + 5 * wd, 5 * wd * sizeof(u64));
    // the cap: nodes[n_shards .. 2 n_shards) are the shard roots; finish nodes[1 .. n_shards) from them
    std::vector<u64> cap(10 * n_shards);
    TF21_TRY(tf21_merkle_build(nodes_out + 5 * n_shards, n_shards, cap.data()));
    std::memcpy(nodes_out, cap.data(), 5 * n_shards * sizeof(u64));  // includes nodes[0] = 0
    return 0;
}

// One shard per device: every device builds its subtree (merkle_tree.rs:247-275), the shard roots -- the tree cap --
// are exchanged with ONE ncclAllGather of 40 bytes per rank over NVLink, and every device hashes the top
// log2(n_shards) levels itself.  Nothing but the cap crosses devices.
static int merkle_build_sharded_nccl(const uint64_t *leafs, uint64_t n_leafs, uint64_t *nodes_out, uint32_t n_shards) {
    tf21_ncclComm_t *comms = nullptr;
    TF21_TRY(nccl_comms((int)n_shards, &comms));
    const u64 local = n_leafs / n_shards;
    struct Shard {
        cudaStream_t st = nullptr;
        u64 *d_leafs = nullptr, *d_nodes = nullptr, *d_roots = nullptr, *d_cap = nullptr;
        int rc = 0;
    };
    std::vector<Shard> sh(n_shards);
    int prev = 0;
    cudaGetDevice(&prev);
    auto on_all = [&](auto &&fn) {
        std::vector<std::thread> workers;
        for (uint32_t i = 0; i < n_shards; i++)
            workers.emplace_back([&, i] {
                if (sh[i].rc) return;
                if (cudaSetDevice((int)i) != cudaSuccess) {
                    sh[i].rc = cuda_fail(cudaGetLastError(), "cudaSetDevice", __LINE__);
                    return;
                }
                sh[i].rc = fn(i, sh[i]);
            });
        for (auto &w : workers) w.join();
    };
    auto first_error = [&]() {
        for (auto &x : sh)
            if (x.rc) return x.rc;
        return 0;
    };
    // phase 1: local subtrees, device resident
    on_all([&](uint32_t i, Shard &x) -> int {
        TF21_CUDA(cudaStreamCreateWithFlags(&x.st, cudaStreamNonBlocking));
        TF21_CUDA(cudaMalloc((void **)&x.d_leafs, 5 * local * sizeof(u64)));
        TF21_CUDA(cudaMalloc((void **)&x.d_nodes, 10 * local * sizeof(u64)));
        TF21_CUDA(cudaMalloc((void **)&x.d_roots, 5 * n_shards * sizeof(u64)));
        TF21_CUDA(cudaMalloc((void **)&x.d_cap, 10 * n_shards * sizeof(u64)));
        TF21_CUDA(cudaMemcpyAsync(x.d_leafs, leafs + 5 * i * local, 5 * local * sizeof(u64), cudaMemcpyHostToDevice, x.st));
        TF21_TRY(tf21_merkle_build_dev(x.d_leafs, local, x.d_nodes, (tf21_stream_t)x.st));
        TF21_CUDA(cudaStreamSynchronize(x.st));
        return 0;
    });
    int rc = first_error();
    // phase 2: the cap -- one all-gather of the shard roots (nodes[1] of every local tree), launched as one group
    if (!rc) {
        int nrc = g_nccl.GroupStart();
        for (uint32_t i = 0; i < n_shards && nrc == 0; i++)
            nrc = g_nccl.AllGather(sh[i].d_nodes + 5, sh[i].d_roots, 5, kNcclUint64, comms[i], sh[i].st);
        const int erc = g_nccl.GroupEnd();
        if (nrc == 0) nrc = erc;
        if (nrc != 0) rc = nccl_fail(nrc, "ncclAllGather");
    }
    // phase 3: every device finishes the top of the tree from the gathered cap; results go home
    if (!rc) {
        on_all([&](uint32_t i, Shard &x) -> int {
            TF21_TRY(tf21_merkle_build_dev(x.d_roots, n_shards, x.d_cap, (tf21_stream_t)x.st));
            for (u64 wd = 1; wd <= local; wd <<= 1)
                TF21_CUDA(cudaMemcpyAsync(nodes_out + 5 * (n_shards * wd + i * wd), x.d_nodes + 5 * wd,
                                          5 * wd * sizeof(u64), cudaMemcpyDeviceToHost, x.st));
            if (i == 0)  // nodes[0] = 0 and nodes[1 .. n_shards): identical on every device, device 0 reports them
                TF21_CUDA(cudaMemcpyAsync(nodes_out, x.d_cap, 5 * n_shards * sizeof(u64), cudaMemcpyDeviceToHost, x.st));
            TF21_CUDA(cudaStreamSynchronize(x.st));
            return 0;
        });
        rc = first_error();
    }
    for (uint32_t i = 0; i < n_shards; i++) {
        cudaSetDevice((int)i);
        if (sh[i].st) cudaStreamSynchronize(sh[i].st);
        cudaFree(sh[i].d_leafs);
        cudaFree(sh[i].d_nodes);
        cudaFree(sh[i].d_roots);
        cudaFree(sh[i].d_cap);
        if (sh[i].st) cudaStreamDestroy(sh[i].st);
    }
    cudaSetDevice(prev);
    return rc;
}

int tf21_merkle_build_sharded(const uint64_t *leafs, uint64_t n_leafs, uint64_t *nodes_out, uint32_t n_shards) {
    TF21_TRY(check_leaf_count(n_leafs));
    if (!leafs || !nodes_out) return TF21_E_BAD_ARG;
    if (n_shards == 0 || (n_shards & (n_shards - 1))) return TF21_E_BAD_ARG;  // subtrees are powers of two
    while (n_shards > 1 && n_leafs / n_shards < 1) n_shards >>= 1;
    if (n_shards == 1) return tf21_merkle_build(leafs, n_leafs, nodes_out);
    int n_dev = 0;
    TF21_TRY(visible_devices(&n_dev));
    // one shard per device: the cap travels over NCCL.  More shards than devices (index algebra testable on one
    // GPU), TF21_NO_NCCL, or no libnccl at all: the cap is assembled in host memory.
    if ((int)n_shards <= n_dev && !getenv("TF21_NO_NCCL")) {
        const int rc = merkle_build_sharded_nccl(leafs, n_leafs, nodes_out, n_shards);
        bool have_nccl;
        {
            std::lock_guard<std::mutex> lock(g_nccl_mutex);
            have_nccl = g_nccl.handle != nullptr;
        }
        if (rc != TF21_E_NCCL || have_nccl) return rc;  // a real NCCL failure is reported, a missing library is not
    }
    return merkle_build_sharded_host(leafs, n_leafs, nodes_out, n_shards, n_dev);
}

/* which path tf21_merkle_build_sharded takes for this shard count: 1 = NCCL cap gather, 0 = host-memory cap */
int tf21_sharded_uses_nccl(uint32_t n_shards) {
    int n_dev = 0;
    if (visible_devices(&n_dev) != 0 || n_shards < 2 || (int)n_shards > n_dev || getenv("TF21_NO_NCCL")) return 0;
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    return nccl_load() ? 1 : 0;
}

int tf21_merkle_build(const uint64_t *leafs, uint64_t n_leafs, uint64_t *nodes_out) {
    TF21_TRY(check_leaf_count(n_leafs));
    if (!leafs || !nodes_out) return TF21_E_BAD_ARG;
    DevBuf bl, bn;
    TF21_TRY(bl.alloc(5 * n_leafs));
    TF21_TRY(bn.alloc(10 * n_leafs));
    TF21_TRY(copy_h2d(bl.p, leafs, 5 * n_leafs * sizeof(u64), nullptr));
    TF21_TRY(tf21_merkle_build_dev(bl.p, n_leafs, bn.p, nullptr));
    // nodes[n .. 2n) is a copy of the caller's own leaves (merkle_tree.rs:426): it is written on the host while the
    // device hashes, and only the inner nodes cross PCIe
    std::thread leaf_copy;
    const bool overlap = (leafs + 5 * n_leafs <= nodes_out || nodes_out + 10 * n_leafs <= leafs) && n_leafs >= (1u << 16);
    bool started = false;
    if (overlap) {
        try {
            leaf_copy = std::thread([=] { parallel_memcpy(nodes_out + 5 * n_leafs, leafs, 5 * n_leafs * sizeof(u64)); });
            started = true;
        } catch (...) {
        }
    }
    if (!started) memmove(nodes_out + 5 * n_leafs, leafs, 5 * n_leafs * sizeof(u64));
    const int rc = copy_d2h(nodes_out, bn.p, 5 * n_leafs * sizeof(u64), nullptr);
    if (leaf_copy.joinable()) leaf_copy.join();
    return rc;
}

int tf21_merkle_root(const uint64_t *leafs, uint64_t n_leafs, uint64_t root_out[5]) {
    TF21_TRY(check_leaf_count(n_leafs));
    if (!leafs || !root_out) return TF21_E_BAD_ARG;
    DevBuf bl, br;
    TF21_TRY(bl.alloc(5 * n_leafs));
    TF21_TRY(br.alloc(5));
    TF21_TRY(copy_h2d(bl.p, leafs, 5 * n_leafs * sizeof(u64), nullptr));
    TF21_TRY(tf21_merkle_root_dev(bl.p, n_leafs, br.p, nullptr));
    TF21_TRY(copy_d2h(root_out, br.p, 5 * sizeof(u64), nullptr));
    return 0;
}

}  // extern "C"
