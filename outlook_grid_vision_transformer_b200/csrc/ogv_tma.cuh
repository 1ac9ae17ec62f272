// Host-side TMA tensor-map construction shared by the kernels that stage tiles with
// cp.async.bulk.tensor (GEMM operands, depthwise / outlook halo tiles).  The driver entry point is
// resolved at run time (cudaGetDriverEntryPoint), so the library has no link-time libcuda dependency.
#pragma once
#include <cuda.h>
#include "ogv_common.cuh"

// rank-N tiled tensor map over a dense-or-strided tensor of OGV_F32 / OGV_BF16 elements.
//   dims[0] is the contiguous (innermost) extent; strides_bytes[i] is the byte stride of dims[i+1]
//   (rank-1 entries, each a multiple of 16); box[i] <= 256 elements; out-of-bounds elements read as 0.
//   swizzle: 0 none, 1 32B, 2 64B, 3 128B.
int ogv_make_tmap(CUtensorMap* tm, const void* ptr, int dtype, int rank, const unsigned long long* dims,
                  const unsigned long long* strides_bytes, const unsigned* box, int swizzle);

// Exact x / d for 0 <= x < 65536, 1 <= d < 65536 with one IMAD.HI (m = floor(2^32 / d) + 1).
struct FastDiv {
  unsigned m;
  int d;
};
static inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.m = d <= 1 ? 0u : (unsigned)(0x100000000ull / (unsigned long long)d) + 1u;
  return f;
}
__device__ __forceinline__ int fdiv(int x, const FastDiv& f) { return f.d <= 1 ? x : (int)__umulhi((unsigned)x, f.m); }
