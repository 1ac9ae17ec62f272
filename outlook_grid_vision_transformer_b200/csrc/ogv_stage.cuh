// Staging-tile helpers shared by the tcgen05 kernels (GEMM engine, fused MLP): a warp's 32 x 32 bf16 tile in the
// TMA SWIZZLE_64B layout, and row writes into 128B-swizzled K-major operand tiles.
#pragma once
#include "ogv_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// staged epilogue helpers: a warp's 32 x 32 bf16 tile in the TMA SWIZZLE_64B layout
// (byte address bits [4,6) ^= bits [7,9)): conflict-free 16-byte row-owner accesses.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw64(int row, int c16) { return row * 64 + ((c16 ^ ((row >> 1) & 3)) << 4); }

__device__ __forceinline__ void stage_write_row(uint8_t* slot, int row, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[8 * c + 2 * i], v[8 * c + 2 * i + 1]);
    *reinterpret_cast<uint4*>(slot + sw64(row, c)) = u;
  }
}
// packed-pair forms (element pairs (2i, 2i+1) in one 64-bit register, see ogv_common.cuh)
__device__ __forceinline__ void stage_write_row(uint8_t* slot, int row, const f32x2 (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float lo, hi;
      unpk2(v[4 * c + i], lo, hi);
      h[i] = __floats2bfloat162_rn(lo, hi);
    }
    *reinterpret_cast<uint4*>(slot + sw64(row, c)) = u;
  }
}
__device__ __forceinline__ void stage_read_row(const uint8_t* slot, int row, f32x2 (&r)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u = *reinterpret_cast<const uint4*>(slot + sw64(row, c));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      r[4 * c + i] = pk2(f.x, f.y);
    }
  }
}
__device__ __forceinline__ void stage_read_row(const uint8_t* slot, int row, float (&r)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u = *reinterpret_cast<const uint4*>(slot + sw64(row, c));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      r[8 * c + 2 * i] = f.x;
      r[8 * c + 2 * i + 1] = f.y;
    }
  }
}


// Byte offset of the 16-byte chunk `c16` (0..7) of row `row` inside a [rows x 64 bf16] sub-tile stored in the TMA /
// UMMA SWIZZLE_128B K-major layout (rows of 128 B, the chunk index XOR-ed with row % 8; tile base 1024-byte aligned).
__device__ __forceinline__ uint32_t sw128(int row, int c16) { return row * 128 + ((c16 ^ (row & 7)) << 4); }

// 8 consecutive fp32 values (as 4 packed pairs) -> one 16-byte chunk of bf16
__device__ __forceinline__ uint4 pack8_bf16(const f32x2 (&v)[4]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    unpk2(v[i], lo, hi);
    h[i] = __floats2bfloat162_rn(lo, hi);
  }
  return u;
}

}  // namespace
