// tests/emu/include/cuda.h -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
// The slice of the driver API the product uses: tensor maps for TMA tile loads, emulated together with the
// mbarrier transaction counting they complete on.
#pragma once
#include "cuda_runtime.h"

typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

struct __attribute__((aligned(64))) CUtensorMap {
    uint64_t opaque[16];  // emulator's own encoding: base, dim0, dim1, stride1, box0, box1, swizzle
};

namespace emu {

inline CUresult encode_tiled(CUtensorMap *map, CUtensorMapDataType dt, cuuint32_t rank, void *base, const cuuint64_t *dims,
                             const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *estr,
                             CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (dt != CU_TENSOR_MAP_DATA_TYPE_UINT8 || rank != 2 || il != CU_TENSOR_MAP_INTERLEAVE_NONE) return CUDA_ERROR_INVALID_VALUE;
    if (estr[0] != 1 || estr[1] != 1) return CUDA_ERROR_INVALID_VALUE;
    if (((uintptr_t)base & 15) || (strides[0] & 15)) return CUDA_ERROR_INVALID_VALUE;  // the driver's alignment rules
    if (box[0] > 256 || box[1] > 256 || box[0] == 0 || box[1] == 0) return CUDA_ERROR_INVALID_VALUE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && box[0] > 128) return CUDA_ERROR_INVALID_VALUE;
    if (sw != CU_TENSOR_MAP_SWIZZLE_128B && sw != CU_TENSOR_MAP_SWIZZLE_NONE) return CUDA_ERROR_INVALID_VALUE;
    memset(map, 0, sizeof(*map));
    map->opaque[0] = (uint64_t)(uintptr_t)base;
    map->opaque[1] = dims[0];
    map->opaque[2] = dims[1];
    map->opaque[3] = strides[0];
    map->opaque[4] = box[0];
    map->opaque[5] = box[1];
    map->opaque[6] = (uint64_t)sw;
    return CUDA_SUCCESS;
}

// mbarrier word: low half = completed phases, high half = transaction bytes still expected in the current phase
EMU_NOTSAN inline void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }
EMU_NOTSAN inline void mbar_tx(uint64_t *bar, int64_t delta, bool arrive) {
    int64_t pending = (int64_t)(int32_t)(*bar >> 32) + delta;
    uint32_t phases = (uint32_t)*bar;
    (void)arrive;
    if (pending == 0) {  // the single expected arrival has happened and every byte has landed
        phases++;
        EMU_TSAN(__tsan_release(bar);)  // phase completion publishes the tile to whoever waits on the barrier
    }
    *bar = ((uint64_t)(uint32_t)(int32_t)pending << 32) | phases;
}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { mbar_tx(bar, (int64_t)bytes, true); }
EMU_NOTSAN inline void mbar_wait(uint64_t *bar, uint32_t parity) {
    while ((((uint32_t)*reinterpret_cast<volatile uint64_t *>(bar)) & 1u) == parity) yield();
    EMU_TSAN(__tsan_acquire(bar);)
}
// cp.async.bulk.tensor.2d: box (box0 x box1) at element coordinates (c0, c1); out-of-bounds elements read as zero;
// SWIZZLE_128B: within each 128-byte row of the destination, 16-byte chunk j lands at chunk j ^ (row % 8), rows
// counted from the (1024-byte aligned) destination.
inline void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    const uint8_t *base = (const uint8_t *)(uintptr_t)map->opaque[0];
    const int64_t dim0 = (int64_t)map->opaque[1], dim1 = (int64_t)map->opaque[2], stride = (int64_t)map->opaque[3];
    const int box0 = (int)map->opaque[4], box1 = (int)map->opaque[5];
    const bool swz = map->opaque[6] == (uint64_t)CU_TENSOR_MAP_SWIZZLE_128B;
    uint8_t *d = (uint8_t *)dst;
    if (swz && (((uintptr_t)d - (uintptr_t)smem_base()) & 127)) die("TMA: swizzled destination not 128-byte aligned");
    for (int r = 0; r < box1; r++) {
        const int64_t gr = (int64_t)c1 + r;
        for (int c = 0; c < box0; c++) {
            const int64_t gc = (int64_t)c0 + c;
            uint8_t v = 0;
            if (gr >= 0 && gr < dim1 && gc >= 0 && gc < dim0) v = base[gr * stride + gc];
            size_t off = (size_t)r * box0 + c;
            if (swz) {
                // pattern repeats every 1024 bytes of SHARED ADDRESS: row index taken from the absolute offset
                const size_t abs_off = ((uintptr_t)d - (uintptr_t)smem_base()) + off;
                const size_t row = (abs_off >> 7) & 7, chunk = (abs_off >> 4) & 7;
                const size_t sw_off = (abs_off & ~(size_t)0x70) | ((chunk ^ row) << 4);
                smem_base()[sw_off] = v;
                continue;
            }
            d[off] = v;
        }
    }
    mbar_tx(bar, -(int64_t)box0 * box1, false);
}

}  // namespace emu

inline cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
    if (strcmp(name, "cuTensorMapEncodeTiled") == 0) {
        *fn = (void *)&emu::encode_tiled;
        if (q) *q = cudaDriverEntryPointSuccess;
    } else {
        *fn = nullptr;
        if (q) *q = cudaDriverEntryPointSymbolNotFound;
    }
    return cudaSuccess;
}
