// Host-side creation of TMA tensor maps through the driver entry point (no link-time
// dependency on libcuda: the library must load on a box without a GPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

namespace sn {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// float32 grid [B,Z,X,Y] (Y fastest) with a (1, boxZ, boxX, boxY) box; out-of-bounds elements
// read as zero, which is exactly the conv's 'same' zero padding.  Returns false when TMA
// cannot express the tensor (caller falls back to plain loads).
// elem_bytes: 4 (float32) or 8 (float64 grids: the pred / dpred boxes of stencil_bwd_fused.cu)
inline bool make_grid_tmap(CUtensorMap* m, const void* x, int B, int Z, int X, int Y, int boxZ, int boxX, int boxY,
                           int elem_bytes = 4) {
    memset(m, 0, sizeof(*m));
    static const bool disabled = SN_ENV("SN_NO_TMA") != nullptr;  // debugging aid: force the plain-load path
    if (disabled) return false;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const int per16 = 16 / elem_bytes;
    if ((Y % per16) || ((uintptr_t)x & 15)) return false;           // global strides must be multiples of 16 B
    if (boxY > 256 || boxX > 256 || boxZ > 256 || (boxY % per16)) return false;
    const cuuint64_t eb = (cuuint64_t)elem_bytes;
    cuuint64_t dims[4] = {(cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)Z, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Y * eb, (cuuint64_t)X * Y * eb, (cuuint64_t)Z * X * Y * eb};
    cuuint32_t box[4] = {(cuuint32_t)boxY, (cuuint32_t)boxX, (cuuint32_t)boxZ, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace sn
