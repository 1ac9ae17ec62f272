// Column-reduction skeleton for "rows x channels" tensors: every thread owns one 8-channel vector
// column (cv) and strides over rows, so loads are fully coalesced and the per-channel partial sums
// live in registers; one shared-memory fold per CTA and one atomicAdd per channel per CTA.
#pragma once
#include "ogv_common.cuh"

struct ColReduceCfg {
  int grid, block;
};

// nv = channel vectors (8 channels each) per row; block = nv * k threads with k = floor(T / nv),
// T = COLREDUCE_THREADS.  Few, fat CTAs (2 per SM): every CTA ends with one atomicAdd per channel on
// the SAME few cache lines, and same-line atomics serialise in L2 -- with 8 CTAs/SM that tail cost
// more than the streaming pass itself.
constexpr int COLREDUCE_THREADS = 512;
static inline bool colreduce_config(long long M, int nv, ColReduceCfg* cfg) {
  if (nv < 1 || nv > 256) return false;
  int k = COLREDUCE_THREADS / nv;
  cfg->block = nv * k;
  long long blocks = (M + k - 1) / k;
  long long cap = (long long)ogv_num_sms() * 2;
  cfg->grid = (int)(blocks < cap ? blocks : cap);
  if (cfg->grid < 1) cfg->grid = 1;
  return true;
}

#define COLREDUCE_LOOP(M, nv, row, cv)                                                \
  const int cv = threadIdx.x % (nv);                                                  \
  const long long _cr_stride = (long long)gridDim.x * (blockDim.x / (nv));            \
  _Pragma("unroll 4")                                                                 \
  for (long long row = (long long)blockIdx.x * (blockDim.x / (nv)) + threadIdx.x / (nv); row < (M); \
       row += _cr_stride)

template <int Q>
__device__ __forceinline__ void colreduce_init(float (&acc)[Q][8]) {
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
}

// outs[q][c] += column sums; every thread of the CTA must call this (it synchronises).
template <int Q>
__device__ __forceinline__ void colreduce_finish(float (&acc)[Q][8], float* const (&outs)[Q], int nv) {
  __shared__ float red[COLREDUCE_THREADS * 9];
  const int k = blockDim.x / nv;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 9 + i] = acc[q][i];
    __syncthreads();
    if (threadIdx.x < nv && outs[q] != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = 0.f;
        for (int j = 0; j < k; ++j) s += red[(threadIdx.x + j * nv) * 9 + i];
        atomicAdd(outs[q] + threadIdx.x * 8 + i, s);
      }
    }
    __syncthreads();
  }
}
