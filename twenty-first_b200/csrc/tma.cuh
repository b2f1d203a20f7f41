// tma.cuh -- Tensor Memory Accelerator plumbing for the tile-moving passes (sm_100a).
//
// The strided side of a 1024-point pass is a [1024 rows][4 word-columns] tile whose rows are 32 bytes wide and
// `inner_words * 8` bytes apart in HBM.  Instead of 32 per-lane 8-byte copies per thread in each direction
// (LDGSTS / STG with their address arithmetic in the ALU-bound instruction stream), one elected thread issues
// four `cp.async.bulk.tensor.3d` box copies of 256 rows each (UTMALDG / UTMASTG in SASS); completion is signalled
// on an mbarrier (loads) or through the bulk async-group (stores).  The tensor map is encoded on the host per
// call (cuTensorMapEncodeTiled, resolved through cudaGetDriverEntryPoint so that libcuda is not a link-time
// dependency) and passed as a __grid_constant__ kernel parameter.
//
// Shared-memory layout of a tile as TMA writes it with CU_TENSOR_MAP_SWIZZLE_32B: row r at byte r * 32, the two
// 16-byte halves of the row swapped when bit 7 of the address (= bit 2 of r) is set.  Word (r, c) of the tile is
// therefore at u64 index  4 r + 2 ((c >> 1) ^ ((r >> 2) & 1)) + (c & 1): lanes that walk down a column (r = 32 a +
// lane) touch 8 of the 16 bank pairs instead of 4, i.e. a column read or write costs 4 wavefronts (2 is ideal).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "runtime.cuh"

namespace tf21 {

typedef CUresult (*tf21_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill);

inline tf21_encode_tiled_fn tma_encoder() {
    static tf21_encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (tf21_encode_tiled_fn)p;
    }();
    return fn;
}

constexpr u32 kTmaTileCols = 4;     // word-columns per tile (32-byte rows: one DRAM sector)
constexpr u32 kTmaBoxRows = 256;    // rows per box copy (the hardware limit per dimension)

// [slabs][1024 rows][inner_words] u64 tensor at `base` (16-byte aligned), boxes of 256 rows x 4 words.
// Returns false when TMA cannot describe it (alignment, no driver entry point): the caller takes the LDGSTS path.
inline bool tma_encode_tile_map(CUtensorMap *map, const u64 *base, u64 inner_words, u64 slabs) {
    tf21_encode_tiled_fn enc = tma_encoder();
    if (!enc || ((uintptr_t)base & 15) != 0 || (inner_words & 1) != 0 || slabs == 0 || slabs > 0xffffffffull ||
        inner_words > 0xffffffffull)
        return false;
    const cuuint64_t dims[3] = {inner_words, 1024, slabs};
    const cuuint64_t strides[2] = {inner_words * 8, 1024 * inner_words * 8};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {kTmaTileCols, kTmaBoxRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (strides[1] >= (1ull << 40)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<u64 *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TF21_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra TF21_MBAR_WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared box copy, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, u32 c0, u32 c1, u32 c2, u64 *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global box copy, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, u32 c0, u32 c1, u32 c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the source tile may be reused (or the CTA may exit) once the copies have READ shared memory
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// u64 index of word (row r, column c) of a [1024][4] tile in the SWIZZLE_32B layout (tile base 256-byte aligned)
__device__ __forceinline__ u32 tma_tile_word(u32 r, u32 c) { return 4u * r + 2u * ((c >> 1) ^ ((r >> 2) & 1u)) + (c & 1u); }
#endif

}  // namespace tf21
