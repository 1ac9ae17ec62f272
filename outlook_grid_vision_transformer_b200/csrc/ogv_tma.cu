#include "ogv_tma.cuh"

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
}  // namespace

int ogv_make_tmap(CUtensorMap* tm, const void* ptr, int dtype, int rank, const unsigned long long* dims,
                  const unsigned long long* strides_bytes, const unsigned* box, int swizzle) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    ogv_set_error("cuTensorMapEncodeTiled entry point not available");
    return OGV_ERR_CUDA;
  }
  if (rank < 1 || rank > 5) {
    ogv_set_error("tensor map rank %d out of range", rank);
    return OGV_ERR_ARG;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1u;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  const CUtensorMapDataType dt = dtype == OGV_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = swizzle == 3   ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                               : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(tm, dt, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ogv_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,.. box %u,%u,.. ptr %p", (int)r, rank,
                  dims[0], rank > 1 ? dims[1] : 0ull, box[0], rank > 1 ? box[1] : 0u, ptr);
    return OGV_ERR_CUDA;
  }
  return OGV_OK;
}
