// Column-reduction skeleton for "rows x channels" tensors: every thread owns one 8-channel vector
// column (cv) and strides over rows, so loads are fully coalesced and the per-channel partial sums
// live in registers; one shared-memory fold per CTA and one atomicAdd per channel per CTA.
#pragma once
#include <cmath>
#include "ogv_common.cuh"

struct ColReduceCfg {
  int grid, block;
};

// nv = channel vectors (8 channels each) per row; block = nv * k threads with k = floor(T / nv),
// T = COLREDUCE_THREADS.  Grid: every CTA ends with one atomicAdd per channel on the SAME few cache lines and
// those serialise in L2 (measured ~30 G atomics/s), so the tail grows with CTAs x C while the streaming part
// wants every resident slot filled (4 CTAs/SM reach the copy rate on the stage-0 tensors, 2 do not).  Balancing
// the two gives grid ~ sqrt(4 M), clamped to one resident wave (occupancy of the calling kernel x SMs).
constexpr int COLREDUCE_THREADS = 512;
template <typename K>
static inline int colreduce_occupancy(K kernel, int block) {
  static int cache[COLREDUCE_THREADS + 1] = {0};  // one table per kernel instantiation
  if (cache[block] == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0) != cudaSuccess || occ < 1) occ = 1;
    cache[block] = occ > 4 ? 4 : occ;
  }
  return cache[block];
}
template <typename K>
static inline bool colreduce_config(long long M, int nv, ColReduceCfg* cfg, K kernel) {
  if (nv < 1 || nv > 256) return false;
  int k = COLREDUCE_THREADS / nv;
  cfg->block = nv * k;
  long long blocks = (M + k - 1) / k;
  long long cap = (long long)ogv_num_sms() * colreduce_occupancy(kernel, cfg->block);
  long long want = (long long)sqrt(4.0 * (double)M) + 1;
  if (want < 64) want = 64;
  if (cap > want) cap = want;
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
// Narrow rows (nv divides 32: C = 64 / 128 / 256) first fold the 32 / nv lanes of a warp that share a channel
// vector with shuffles; the per-warp (or per-thread-row) partials are then summed by C threads in parallel, one
// channel each -- the fold used to run on nv threads x 8 channels x k rows serially (512 dependent shared-memory
// reads at C = 64, a third of the run time of a 30 us reduction).
template <int Q>
__device__ __forceinline__ void colreduce_finish(float (&acc)[Q][8], float* const (&outs)[Q], int nv) {
  __shared__ float red[COLREDUCE_THREADS * 9];
  const bool warp_fold = nv < 32 && (32 % nv) == 0;  // then blockDim.x == COLREDUCE_THREADS, whole warps only
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = warp_fold ? (int)(blockDim.x >> 5) : (int)blockDim.x / nv;
  const int C = nv * 8;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    if (warp_fold) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = acc[q][i];
        for (int o = nv; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < nv) red[(warp * nv + lane) * 9 + i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[threadIdx.x * 9 + i] = acc[q][i];
    }
    __syncthreads();
    if (outs[q] != nullptr) {
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int cv = c >> 3, i = c & 7;
        float s = 0.f;
        for (int j = 0; j < rows; ++j) s += red[(cv + j * nv) * 9 + i];
        atomicAdd(outs[q] + c, s);
      }
    }
    __syncthreads();
  }
}
