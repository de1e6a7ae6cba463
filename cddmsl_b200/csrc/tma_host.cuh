// Host-side helper shared by the kernels that use TMA tensor maps: looks cuTensorMapEncodeTiled up through the
// runtime (no -lcuda at link time) and encodes a 2-D fp32 tiled map.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace cddmsl {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn tma_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp32 [dim1 rows][dim0 cols], `row_bytes` between rows; box = [box1 rows][box0 cols].  Returns 0 on success.
static inline int tma_encode_2d_f32(CUtensorMap* m, const void* base, unsigned long long dim0, unsigned long long dim1,
                                    unsigned long long row_bytes, unsigned box0, unsigned box1, CUtensorMapSwizzle swz,
                                    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE) {
  EncodeTiledFn enc = tma_get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

// fp32 [dim2][dim1][dim0], byte strides of dim1 / dim2; box = [box2][box1][box0], no swizzle.  Returns 0 on success.
static inline int tma_encode_3d_f32(CUtensorMap* m, const void* base, unsigned long long dim0, unsigned long long dim1,
                                    unsigned long long dim2, unsigned long long stride1_bytes,
                                    unsigned long long stride2_bytes, unsigned box0, unsigned box1, unsigned box2) {
  EncodeTiledFn enc = tma_get_encode();
  if (!enc) return -1;
  cuuint64_t dims[3] = {(cuuint64_t)dim0, (cuuint64_t)dim1, (cuuint64_t)dim2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1_bytes, (cuuint64_t)stride2_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace cddmsl
