// Fused MLP for sm_100a:  y = residual + row_scale * ( act(x W1^T + b1) W2^T + b2 )   in ONE kernel.
// (reference: MLP2d, outlook_attention.py:43-49; MLP, Out_Grid_Block.py:24-32; the residual / DropPath of
//  Outlook_Block.py:63 and Out_Grid_Block.py:102-104)
//
// The [M, hidden] activation never exists in HBM.  Per 128-row tile the hidden dimension is walked in chunks of BH
// columns:   fc1 chunk  Z_j = X W1_j^T      tcgen05.mma  -> TMEM (double buffered)
//            epilogue   H_j = act(Z_j + b1) tcgen05.ld -> registers -> bf16 -> 128B-swizzled K-major smem tile
//            fc2 chunk  Y  += H_j W2_j^T    tcgen05.mma with the smem tile as its A operand -> TMEM
// and after the last chunk  y = (Y + b2) * row_scale + residual  leaves through swizzled staging + TMA store.
// Warp roles as in the GEMM engine: warps 0..15 epilogue (lane = row, warp % 4 = TMEM lane quarter, warp / 4 = column
// group), warp 16 TMA producer (x tile, weight-chunk ring), warp 17 single-thread MMA issuer.  The MMA issuer runs
// fc1 of chunk j+1 BEFORE fc2 of chunk j, so the tensor pipe works while the epilogue warps compute the activation.
#include <stdlib.h>

#include "ogv_gemm.cuh"
#include "ogv_ptx.cuh"
#include "ogv_stage.cuh"
#include "ogv_tma.cuh"

namespace {

constexpr int F_BM = 128;
constexpr int F_EPI_WARPS = 16;
constexpr int F_THREADS = (F_EPI_WARPS + 2) * 32;
constexpr int F_PRODUCER = F_EPI_WARPS;
constexpr int F_MMA = F_EPI_WARPS + 1;
constexpr int F_SLOT = 32 * 32 * 2;     // one warp's 32 x 32 bf16 staging tile (SWIZZLE_64B)

template <int C, int BH>
struct MlpFwdCfg {
  static_assert(C % 64 == 0 && BH % 64 == 0, "tiles are cut in 64-column (128-byte) sub-tiles");
  static constexpr int KC = C / 64;
  static constexpr int KH = BH / 64;
  static constexpr int NW = C == 64 ? 4 : 3;               // weight-chunk ring depth
  static constexpr int NZ = (512 - 2 * C) / BH > 6 ? 6 : (512 - 2 * C) / BH;  // fc1 accumulator buffers in TMEM
  static constexpr int NHB = (BH == 64 && C == 64) ? 2 : 1;  // hidden-tile buffers per epilogue group
  static constexpr int XA_BYTES = F_BM * C * 2;
  static constexpr int H_BYTES = F_BM * BH * 2;
  static constexpr int W_BYTES = BH * C * 2;
  static constexpr int NOC = (C / 32) / 2;                 // 32-column output chunks per warp of a group (1 or 2)
  static constexpr int STAGE_BYTES = F_EPI_WARPS * NOC * F_SLOT;
  static constexpr int NBAR = 2 + 2 + 2 * NW + 2 * NZ + 4 * NHB + 2 + 2 + F_EPI_WARPS * NOC;
  static constexpr int SMEM = 2 * XA_BYTES + 2 * NHB * H_BYTES + NW * W_BYTES + STAGE_BYTES + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 512;                    // NZ*BH (Z) + 2*C (Y) <= 512
  static_assert(NZ * BH + 2 * C <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct MlpFwdParams {
  long long M;
  int Hd, NJ, m_tiles;
  int act;
  int has_res;
  const float* b1;
  const float* b2;
  const float* row_scale;
  int rows_per_scale;
};

// The kernel sees the work of one CTA as a FLAT STREAM of hidden chunks f = 0, 1, ... (tile it = f / NJ, chunk
// j = f % NJ).  TMEM reads are the scarce resource of the epilogue (64 B/clk/SM: a 128 x 128 fp32 chunk costs 1024
// clocks to read, about as much as its GELU costs to compute), so the 16 epilogue warps form TWO groups that work on
// alternate chunks and drift into anti-phase -- one group reads TMEM while the other one computes.  The MMA issuer runs
// fc1 two chunks ahead of fc2 (three Z buffers), so a group that finishes chunk f finds Z of chunk f + 2 waiting.
template <int C, int BH>
__global__ void __launch_bounds__(F_THREADS, 1)
mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmR, const MlpFwdParams p) {
  using Cfg = MlpFwdCfg<C, BH>;
  constexpr int NW = Cfg::NW, NZ = Cfg::NZ, NHB = Cfg::NHB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xa = smem;                                   // [2][XA_BYTES]
  uint8_t* hbuf = xa + 2 * Cfg::XA_BYTES;               // [2 groups][NHB][H_BYTES]
  uint8_t* wring = hbuf + 2 * NHB * Cfg::H_BYTES;       // [NW][W_BYTES]
  uint8_t* staging = wring + NW * Cfg::W_BYTES;         // [F_EPI_WARPS][NOC][F_SLOT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGE_BYTES);
  uint64_t* xa_full = bars;
  uint64_t* xa_empty = xa_full + 2;
  uint64_t* w_full = xa_empty + 2;
  uint64_t* w_empty = w_full + NW;
  uint64_t* z_full = w_empty + NW;
  uint64_t* z_empty = z_full + NZ;
  uint64_t* h_full = z_empty + NZ;
  uint64_t* h_empty = h_full + 2 * NHB;
  uint64_t* y_full = h_empty + 2 * NHB;
  uint64_t* y_empty = y_full + 2;
  uint64_t* ld_bar = y_empty + 2;                       // [F_EPI_WARPS][NOC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + F_EPI_WARPS * Cfg::NOC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == F_PRODUCER && lane == 0) {
    ptx::tma_prefetch_desc(&tmX);
    ptx::tma_prefetch_desc(&tmW1);
    ptx::tma_prefetch_desc(&tmW2);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&xa_full[s], 1);
      ptx::mbar_init(&xa_empty[s], 1);
      ptx::mbar_init(&y_full[s], 1);
      ptx::mbar_init(&y_empty[s], F_EPI_WARPS / 2);
    }
    for (int s = 0; s < 2 * NHB; ++s) {
      ptx::mbar_init(&h_full[s], F_EPI_WARPS / 2);
      ptx::mbar_init(&h_empty[s], 1);
    }
    for (int s = 0; s < NZ; ++s) {
      ptx::mbar_init(&z_full[s], 1);
      ptx::mbar_init(&z_empty[s], F_EPI_WARPS / 2);
    }
    for (int s = 0; s < NW; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < F_EPI_WARPS * Cfg::NOC; ++s) ptx::mbar_init(&ld_bar[s], 1);
    ptx::fence_barrier_init();
  }
  if (warp == F_MMA) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_z = tmem_base;               // Z[b] at column b * BH
  const uint32_t tmem_y = tmem_base + NZ * BH;     // Y[b] at column NZ*BH + b * C
  const int NJ = p.NJ;
  const int my_tiles = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  const int T = my_tiles * NJ;                     // chunks of this CTA

  if (warp == F_PRODUCER) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t wc = 0;
      auto slot_for = [&]() {
        const int ws = wc % NW;
        ptx::mbar_wait(&w_empty[ws], ((wc / NW) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::W_BYTES);
        ++wc;
        return ws;
      };
      for (int f = 0; f < T + NZ; ++f) {  // same order as the MMA issuer consumes: [x tile], W1(f), W2(f - NZ)
        if (f < T) {
          const int it = f / NJ, j = f - it * NJ;
          if (j == 0) {
            const int xb = it & 1;
            const int tile = blockIdx.x + it * gridDim.x;
            ptx::mbar_wait(&xa_empty[xb], ((it >> 1) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&xa_full[xb], Cfg::XA_BYTES);
#pragma unroll
            for (int ks = 0; ks < Cfg::KC; ++ks)  // box {64 k, 128 rows}
              ptx::tma_load_2d(xa + xb * Cfg::XA_BYTES + ks * (F_BM * 128), &tmX, &xa_full[xb], ks * 64, tile * F_BM);
          }
          const int ws = slot_for();  // W1 rows [j*BH, +BH): box {64 k, BH rows} per 64-column slice of C
#pragma unroll
          for (int ks = 0; ks < Cfg::KC; ++ks)
            ptx::tma_load_2d(wring + ws * Cfg::W_BYTES + ks * (BH * 128), &tmW1, &w_full[ws], ks * 64, j * BH);
        }
        if (f >= NZ) {
          const int j2 = (f - NZ) % NJ;
          const int ws = slot_for();  // W2[:, j2*BH ..): box {64 hidden, C rows} per 64-column slice of the chunk
#pragma unroll
          for (int hs = 0; hs < Cfg::KH; ++hs)
            ptx::tma_load_2d(wring + ws * Cfg::W_BYTES + hs * (C * 128), &tmW2, &w_full[ws], j2 * BH + hs * 64, 0);
        }
      }
    }
  } else if (warp == F_MMA) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc1 = ptx::umma_idesc_bf16(F_BM, BH, 0, 0);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(F_BM, C, 0, 0);
      uint32_t wc = 0;
      // fc1 runs NZ chunks ahead of fc2: Z(f) is issued the moment the epilogue has READ chunk f - NZ out of TMEM, long
      // before that chunk's activation tile is complete, so no epilogue group ever waits for the tensor pipe
      for (int f = 0; f < T + NZ; ++f) {
        if (f < T) {  // fc1 of chunk f
          const int it = f / NJ, j = f - it * NJ;
          const int xb = it & 1;
          if (j == 0) ptx::mbar_wait(&xa_full[xb], (it >> 1) & 1);
          const int zb = f % NZ;
          ptx::mbar_wait(&z_empty[zb], ((f / NZ) & 1) ^ 1u);
          const int ws = wc % NW;
          ptx::mbar_wait(&w_full[ws], (wc / NW) & 1);
          ptx::tc_fence_after();
          const uint32_t xa_addr = ptx::smem_u32(xa + xb * Cfg::XA_BYTES);
          const uint32_t wb = ptx::smem_u32(wring + ws * Cfg::W_BYTES);
#pragma unroll
          for (int ks = 0; ks < Cfg::KC; ++ks)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = ptx::umma_smem_desc(xa_addr + ks * (F_BM * 128) + kk * 32, 16u, 1024u);
              const uint64_t bd = ptx::umma_smem_desc(wb + ks * (BH * 128) + kk * 32, 16u, 1024u);
              ptx::umma_f16(tmem_z + zb * BH, ad, bd, idesc1, (ks > 0 || kk > 0) ? 1u : 0u);
            }
          ptx::umma_commit(&w_empty[ws]);
          ++wc;
          ptx::umma_commit(&z_full[zb]);
          if (j == NJ - 1) ptx::umma_commit(&xa_empty[xb]);
        }
        if (f >= NZ) {  // fc2 of chunk f - NZ
          const int f2 = f - NZ;
          const int it = f2 / NJ, j = f2 - it * NJ;
          const int yb = it & 1;
          const int hb = (f2 & 1) * NHB + ((f2 >> 1) % NHB);  // buffer of group f2 % 2, its (f2 / 2)-th chunk
          if (j == 0) ptx::mbar_wait(&y_empty[yb], ((it >> 1) & 1) ^ 1u);
          ptx::mbar_wait(&h_full[hb], ((f2 >> 1) / NHB) & 1);
          const int ws = wc % NW;
          ptx::mbar_wait(&w_full[ws], (wc / NW) & 1);
          ptx::tc_fence_after();
          const uint32_t ha = ptx::smem_u32(hbuf + hb * Cfg::H_BYTES);
          const uint32_t wb = ptx::smem_u32(wring + ws * Cfg::W_BYTES);
#pragma unroll
          for (int hs = 0; hs < Cfg::KH; ++hs)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = ptx::umma_smem_desc(ha + hs * (F_BM * 128) + kk * 32, 16u, 1024u);
              const uint64_t bd = ptx::umma_smem_desc(wb + hs * (C * 128) + kk * 32, 16u, 1024u);
              ptx::umma_f16(tmem_y + yb * C, ad, bd, idesc2, (j > 0 || hs > 0 || kk > 0) ? 1u : 0u);
            }
          ptx::umma_commit(&w_empty[ws]);
          ++wc;
          ptx::umma_commit(&h_empty[hb]);
          if (j == NJ - 1) ptx::umma_commit(&y_full[yb]);
        }
      }
    }
  } else {
    // ------------------------------- epilogue warps: two groups on alternate chunks -------------------------------
    constexpr int HW = BH / 2;          // hidden columns of a chunk per warp (64 or 32)
    constexpr int NP = HW / 32;         // ... in 32-column pieces
    constexpr int NOC = Cfg::NOC;
    const int g = warp >> 3;            // group: chunks with f % 2 == g, hidden tile buffer H[g]
    const int q = warp & 3;             // TMEM lane quarter
    const int half = (warp >> 2) & 1;   // column half within the group
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* slots = staging + warp * NOC * F_SLOT;
    uint64_t* my_ld = &ld_bar[warp * NOC];
    const int row = q * 32 + lane;
    uint32_t ldc = 0;
    // writes tile `it` out: y = (Y + b2) * row_scale + residual.  Called one own-chunk AFTER the tile's last chunk was
    // converted, so the fc2 accumulator and the residual tile have long arrived.
    auto write_tile = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int yb = it & 1;
      const int mrow0 = tile * F_BM + q * 32;
      const long long m = (long long)mrow0 + lane;
      ptx::mbar_wait(&y_full[yb], (it >> 1) & 1);
      ptx::tc_fence_after();
      const float rsf = (p.row_scale && m < p.M) ? p.row_scale[m / p.rows_per_scale] : 1.f;
      const f32x2 rs = splat2(rsf);
      if (!p.has_res) {
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
      }
#pragma unroll
      for (int oc = 0; oc < NOC; ++oc) {
        const int ccol = (half * NOC + oc) * 32;
        uint8_t* slot = slots + oc * F_SLOT;
        float v[32];
        ptx::tmem_ld32(tmem_y + lane_off + yb * C + ccol, v);
        if (oc == NOC - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&y_empty[yb]);
        }
        f32x2 v2[16];
        pack_n<16>(v, v2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(p.b2 + ccol) + i);
          v2[2 * i] = add2(v2[2 * i], b4.x);
          v2[2 * i + 1] = add2(v2[2 * i + 1], b4.y);
        }
        if (p.has_res) {
          ptx::mbar_wait(&my_ld[oc], ldc & 1u);
          f32x2 r2[16];
          stage_read_row(slot, lane, r2);
#pragma unroll
          for (int i = 0; i < 16; ++i) v2[i] = fma2(v2[i], rs, r2[i]);
          __syncwarp();  // every lane has read its residual row before anyone overwrites the slot
        } else if (p.row_scale) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v2[i] = mul2(v2[i], rs);
        }
        stage_write_row(slot, lane, v2);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::tma_store_2d(&tmY, slot, ccol, mrow0);
      }
      if (lane == 0) ptx::bulk_commit();
      ++ldc;
    };
    int pending = -1;  // tile whose last chunk this group converted and which it still has to write out
    for (int f = g; f < T; f += 2) {
      const int it = f / NJ, j = f - it * NJ;
      const int zb = f % NZ;
      ptx::mbar_wait(&z_full[zb], (f / NZ) & 1);
      ptx::tc_fence_after();
      const int hbi = g * NHB + ((f >> 1) % NHB);
      uint8_t* hb = hbuf + hbi * Cfg::H_BYTES;
#pragma unroll
      for (int pc = 0; pc < NP; ++pc) {
        const int col0 = half * HW + pc * 32;  // within the chunk
        float v[32];
        ptx::tmem_ld32(tmem_z + lane_off + zb * BH + col0, v);
        if (pc == NP - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&z_empty[zb]);
        }
        f32x2 v2[16];
        pack_n<16>(v, v2);
        const float* b1 = p.b1 + j * BH + col0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(b1) + i);
          v2[2 * i] = add2(v2[2 * i], b4.x);
          v2[2 * i + 1] = add2(v2[2 * i + 1], b4.y);
        }
        act_apply_p<16, true>(p.act, v2);
        // H[g] is rewritten only after the fc2 MMA of this group's previous chunk has read it
        if (pc == 0) ptx::mbar_wait(&h_empty[hbi], (((f >> 1) / NHB) & 1) ^ 1u);
        // row `row` of the chunk, columns [col0, +32): four 16-byte pieces of a 128B-swizzled K-major tile
        uint8_t* ht = hb + (col0 / 64) * (F_BM * 128);
        const int c16_0 = (col0 % 64) / 8;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const f32x2 piece[4] = {v2[4 * c], v2[4 * c + 1], v2[4 * c + 2], v2[4 * c + 3]};
          *reinterpret_cast<uint4*>(ht + sw128(row, c16_0 + c)) = pack8_bf16(piece);
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&h_full[hbi]);
      if (pending >= 0) {
        write_tile(pending);
        pending = -1;
      }
      if (j == NJ - 1) {  // this group writes tile `it` out -- after its next chunk; fetch the residual tile meanwhile
        pending = it;
        if (p.has_res && lane == 0) {
          const int mrow0 = (blockIdx.x + it * gridDim.x) * F_BM + q * 32;
          ptx::bulk_wait_read<0>();  // this warp's previous stores have finished reading the slots
#pragma unroll
          for (int oc = 0; oc < NOC; ++oc) {
            ptx::mbar_arrive_expect_tx(&my_ld[oc], F_SLOT);
            ptx::tma_load_2d(slots + oc * F_SLOT, &tmR, &my_ld[oc], (half * NOC + oc) * 32, mrow0);
          }
        }
      }
    }
    if (pending >= 0) write_tile(pending);
    if (lane == 0) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == F_MMA) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// Backward, same idea: per 128-row tile and 64-wide hidden chunk j
//     Z_j  = XN W1_j^T            (recompute of the fc1 pre-activation; the forward saved nothing wide)
//     DH_j = DY W2[:, j]          (= DY (W2^T)_j^T)
//     epilogue:  (h, h') = act, act'(Z_j + b1);  dz = DH_j * h' * s;  hs = h * s        (s = DropPath row scale)
//     DXN += dz W1_j              (dz tile in swizzled smem as the A operand)
// and dz / hs leave through TMA stores from the very tiles the epilogue wrote (they feed the two weight-gradient
// GEMMs:  dW1 = dz^T XN,  dW2 = DY^T hs,  db1 = colsum(dz)).  One wide write each instead of: h + act' written by the
// forward, act' and dz through the fc2 dgrad epilogue, dz read again by the fc1 dgrad.
// =================================================================================================
constexpr int B_BH = 64;

template <int C>
struct MlpBwdCfg {
  static constexpr int KC = C / 64;
  static constexpr int XB = C == 64 ? 2 : 1;               // x/dy tile buffers (the C=128 budget has room for one)
  static constexpr int NW = C == 64 ? 6 : 3;               // weight-chunk ring depth (3 weight chunks per hidden chunk)
  static constexpr int NZ = 3;                             // Z / DH accumulator buffers
  static constexpr int NDX = C == 64 ? 2 : 1;              // DXN accumulator buffers
  static constexpr int XA_BYTES = F_BM * C * 2;            // one activation tile
  static constexpr int T_BYTES = F_BM * B_BH * 2;          // one dz / hs tile (16 KB)
  static constexpr int W_BYTES = B_BH * C * 2;
  static constexpr int NOC = (C / 32) / 2;                 // 32-column dxn chunks per warp of the output group
  static constexpr int STAGE_BYTES = F_EPI_WARPS * F_SLOT;
  static constexpr int NBAR = 2 * XB + 2 * NW + 2 * NZ + 4 + 2 * NDX;
  static constexpr int SMEM = XB * 2 * XA_BYTES + 4 * T_BYTES + NW * W_BYTES + STAGE_BYTES + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 512;                    // Z[3] + DH[3] (6 x 64) + DXN[NDX] (NDX x C)
  static_assert(2 * NZ * B_BH + NDX * C <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct MlpBwdParams {
  long long M;
  int Hd, NJ, m_tiles;
  int act;
  const float* b1;
  const float* row_scale;
  int rows_per_scale;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Same flat chunk stream / two epilogue groups / two-chunk MMA lookahead as the forward kernel.
template <int C>
__global__ void __launch_bounds__(F_THREADS, 1)
mlp_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2T,
               const __grid_constant__ CUtensorMap tmW1T, const __grid_constant__ CUtensorMap tmDZ,
               const __grid_constant__ CUtensorMap tmHS, const __grid_constant__ CUtensorMap tmDX, const MlpBwdParams p) {
  using Cfg = MlpBwdCfg<C>;
  constexpr int BH = B_BH, NW = Cfg::NW, NZ = Cfg::NZ, XB = Cfg::XB, NDX = Cfg::NDX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xn = smem;                                        // [XB][XA_BYTES]
  uint8_t* gy = xn + XB * Cfg::XA_BYTES;                     // [XB][XA_BYTES]
  uint8_t* dzt = gy + XB * Cfg::XA_BYTES;                    // [2][T_BYTES]
  uint8_t* hst = dzt + 2 * Cfg::T_BYTES;                     // [2][T_BYTES]
  uint8_t* wring = hst + 2 * Cfg::T_BYTES;                   // [NW][W_BYTES]
  uint8_t* staging = wring + NW * Cfg::W_BYTES;              // [F_EPI_WARPS][F_SLOT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGE_BYTES);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + XB;
  uint64_t* w_full = x_empty + XB;
  uint64_t* w_empty = w_full + NW;
  uint64_t* zd_full = w_empty + NW;
  uint64_t* zd_empty = zd_full + NZ;
  uint64_t* dz_full = zd_empty + NZ;
  uint64_t* dz_empty = dz_full + 2;
  uint64_t* dx_full = dz_empty + 2;
  uint64_t* dx_empty = dx_full + NDX;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dx_empty + NDX);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == F_PRODUCER && lane == 0) {
    ptx::tma_prefetch_desc(&tmX);
    ptx::tma_prefetch_desc(&tmG);
    ptx::tma_prefetch_desc(&tmW1);
    ptx::tma_prefetch_desc(&tmW2T);
    ptx::tma_prefetch_desc(&tmW1T);
    for (int s = 0; s < XB; ++s) {
      ptx::mbar_init(&x_full[s], 1);
      ptx::mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < NW; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < NZ; ++s) {
      ptx::mbar_init(&zd_full[s], 1);
      ptx::mbar_init(&zd_empty[s], F_EPI_WARPS / 2);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&dz_full[s], F_EPI_WARPS / 2);
      ptx::mbar_init(&dz_empty[s], 1);
    }
    for (int s = 0; s < NDX; ++s) {
      ptx::mbar_init(&dx_full[s], 1);
      ptx::mbar_init(&dx_empty[s], F_EPI_WARPS / 2);
    }
    ptx::fence_barrier_init();
  }
  if (warp == F_MMA) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_z = tmem_base;                     // Z[b]   at column b * 64
  const uint32_t tmem_dh = tmem_base + NZ * BH;          // DH[b]  at column 192 + b * 64
  const uint32_t tmem_dx = tmem_base + 2 * NZ * BH;      // DXN[b] at column 384 + b * C
  const int NJ = p.NJ;
  const int my_tiles = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int T = my_tiles * NJ;

  if (warp == F_PRODUCER) {
    if (lane == 0) {
      uint32_t wc = 0;
      auto slot_for = [&]() {
        const int ws = wc % NW;
        ptx::mbar_wait(&w_empty[ws], ((wc / NW) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::W_BYTES);
        ++wc;
        return ws;
      };
      for (int f = 0; f < T + NZ; ++f) {  // order of consumption: [x, dy tiles], W1(f), W2T(f), W1T(f - NZ)
        if (f < T) {
          const int it = f / NJ, j = f - it * NJ;
          if (j == 0) {
            const int xb = it % XB;
            const int tile = blockIdx.x + it * gridDim.x;
            ptx::mbar_wait(&x_empty[xb], ((it / XB) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&x_full[xb], 2 * Cfg::XA_BYTES);
#pragma unroll
            for (int ks = 0; ks < Cfg::KC; ++ks) {
              ptx::tma_load_2d(xn + xb * Cfg::XA_BYTES + ks * (F_BM * 128), &tmX, &x_full[xb], ks * 64, tile * F_BM);
              ptx::tma_load_2d(gy + xb * Cfg::XA_BYTES + ks * (F_BM * 128), &tmG, &x_full[xb], ks * 64, tile * F_BM);
            }
          }
          int ws = slot_for();  // W1 rows of chunk j: one box {64 k, 64 rows} per 64-column slice of C
#pragma unroll
          for (int ks = 0; ks < Cfg::KC; ++ks)
            ptx::tma_load_2d(wring + ws * Cfg::W_BYTES + ks * (BH * 128), &tmW1, &w_full[ws], ks * 64, j * BH);
          ws = slot_for();      // (W2^T) rows of chunk j
#pragma unroll
          for (int ks = 0; ks < Cfg::KC; ++ks)
            ptx::tma_load_2d(wring + ws * Cfg::W_BYTES + ks * (BH * 128), &tmW2T, &w_full[ws], ks * 64, j * BH);
        }
        if (f >= NZ) {          // (W1^T)[:, chunk]: one box {64 hidden, C rows}
          const int j2 = (f - NZ) % NJ;
          const int ws = slot_for();
          ptx::tma_load_2d(wring + ws * Cfg::W_BYTES, &tmW1T, &w_full[ws], j2 * BH, 0);
        }
      }
    }
  } else if (warp == F_MMA) {
    if (lane == 0) {
      const uint32_t idesc_h = ptx::umma_idesc_bf16(F_BM, BH, 0, 0);
      const uint32_t idesc_x = ptx::umma_idesc_bf16(F_BM, C, 0, 0);
      uint32_t wc = 0;
      for (int f = 0; f < T + NZ; ++f) {  // Z / DH run NZ chunks ahead of the DXN accumulation (see the forward kernel)
        if (f < T) {
          const int it = f / NJ, j = f - it * NJ;
          const int xb = it % XB;
          if (j == 0) ptx::mbar_wait(&x_full[xb], (it / XB) & 1);
          const int zb = f % NZ;
          ptx::mbar_wait(&zd_empty[zb], ((f / NZ) & 1) ^ 1u);
          const uint32_t xn_addr = ptx::smem_u32(xn + xb * Cfg::XA_BYTES);
          const uint32_t gy_addr = ptx::smem_u32(gy + xb * Cfg::XA_BYTES);
#pragma unroll
          for (int which = 0; which < 2; ++which) {  // 0: Z = XN W1_j^T   1: DH = DY (W2^T)_j^T
            const int ws = wc % NW;
            ptx::mbar_wait(&w_full[ws], (wc / NW) & 1);
            ptx::tc_fence_after();
            const uint32_t wb = ptx::smem_u32(wring + ws * Cfg::W_BYTES);
            const uint32_t a0 = which ? gy_addr : xn_addr;
            const uint32_t d = (which ? tmem_dh : tmem_z) + zb * BH;
#pragma unroll
            for (int ks = 0; ks < Cfg::KC; ++ks)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = ptx::umma_smem_desc(a0 + ks * (F_BM * 128) + kk * 32, 16u, 1024u);
                const uint64_t bd = ptx::umma_smem_desc(wb + ks * (BH * 128) + kk * 32, 16u, 1024u);
                ptx::umma_f16(d, ad, bd, idesc_h, (ks > 0 || kk > 0) ? 1u : 0u);
              }
            ptx::umma_commit(&w_empty[ws]);
            ++wc;
          }
          ptx::umma_commit(&zd_full[zb]);
          if (j == NJ - 1) ptx::umma_commit(&x_empty[xb]);
        }
        if (f >= NZ) {  // DXN += dz W1 for chunk f - NZ
          const int f2 = f - NZ;
          const int it = f2 / NJ, j = f2 - it * NJ;
          const int ob = it % NDX, hb = f2 & 1;
          if (j == 0) ptx::mbar_wait(&dx_empty[ob], ((it / NDX) & 1) ^ 1u);
          ptx::mbar_wait(&dz_full[hb], (f2 >> 1) & 1);
          const int ws = wc % NW;
          ptx::mbar_wait(&w_full[ws], (wc / NW) & 1);
          ptx::tc_fence_after();
          const uint32_t da = ptx::smem_u32(dzt + hb * Cfg::T_BYTES);
          const uint32_t wb = ptx::smem_u32(wring + ws * Cfg::W_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = ptx::umma_smem_desc(da + kk * 32, 16u, 1024u);
            const uint64_t bd = ptx::umma_smem_desc(wb + kk * 32, 16u, 1024u);
            ptx::umma_f16(tmem_dx + ob * C, ad, bd, idesc_x, (j > 0 || kk > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&w_empty[ws]);
          ++wc;
          ptx::umma_commit(&dz_empty[hb]);
          if (j == NJ - 1) ptx::umma_commit(&dx_full[ob]);
        }
      }
    }
  } else {
    // ------------------------------- epilogue warps: two groups on alternate chunks -------------------------------
    constexpr int ZW = 16;              // columns converted at a time (register budget)
    constexpr int NOC = Cfg::NOC;
    const int g = warp >> 3;
    const int q = warp & 3;
    const int half = (warp >> 2) & 1;   // 32 of the chunk's 64 hidden columns
    const bool store_thread = half == 0 && lane == 0;  // issues the (group, quarter) slab stores of dz / hs
    const int bar_id = 1 + g * 4 + q;   // the two warps (half 0 / 1) that share a 32-row slab
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* slot = staging + warp * F_SLOT;
    const int row = q * 32 + lane;
    auto write_tile = [&](int it) {  // dxn tile of `it`, one own-chunk after its last chunk (the accumulator has arrived)
      const int ob = it % NDX;
      const int mrow0 = (blockIdx.x + it * gridDim.x) * F_BM + q * 32;
      ptx::mbar_wait(&dx_full[ob], (it / NDX) & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int oc = 0; oc < NOC; ++oc) {
        const int ccol = (half * NOC + oc) * 32;
        float v[32];
        ptx::tmem_ld32(tmem_dx + lane_off + ob * C + ccol, v);
        if (oc == NOC - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&dx_empty[ob]);
        }
        if (lane == 0) ptx::bulk_wait_read<0>();  // earlier stores of this thread have finished reading the slot
        __syncwarp();
        stage_write_row(slot, lane, v);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tmDX, slot, ccol, mrow0);
          ptx::bulk_commit();
        }
      }
    };
    int pending = -1;
    for (int f = g; f < T; f += 2) {
      const int it = f / NJ, j = f - it * NJ;
      const int mrow0 = (blockIdx.x + it * gridDim.x) * F_BM + q * 32;
      const long long m = (long long)mrow0 + lane;
      const float rsf = p.row_scale ? (m < p.M ? p.row_scale[m / p.rows_per_scale] : 0.f) : 1.f;
      const f32x2 rs = splat2(rsf);
      const int zb = f % NZ;
      ptx::mbar_wait(&zd_full[zb], (f / NZ) & 1);
      ptx::tc_fence_after();
      uint8_t* dzb = dzt + g * Cfg::T_BYTES;
      uint8_t* hsb = hst + g * Cfg::T_BYTES;
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        const int col0 = half * 32 + pc * ZW;
        float zf[ZW], df[ZW];
        ptx::tmem_ld16(tmem_z + lane_off + zb * BH + col0, zf);
        ptx::tmem_ld16(tmem_dh + lane_off + zb * BH + col0, df);
        if (pc == 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&zd_empty[zb]);
        }
        f32x2 z2[ZW / 2], d2[ZW / 2], g2[ZW / 2];
        pack_n<ZW / 2>(zf, z2);
        pack_n<ZW / 2>(df, d2);
        const float* b1 = p.b1 + j * BH + col0;
#pragma unroll
        for (int i = 0; i < ZW / 4; ++i) {
          const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(b1) + i);
          z2[2 * i] = add2(z2[2 * i], b4.x);
          z2[2 * i + 1] = add2(z2[2 * i + 1], b4.y);
        }
        act_both_p<ZW / 2, true>(p.act, z2, g2);  // z2 <- act(z), g2 <- act'(z)
#pragma unroll
        for (int i = 0; i < ZW / 2; ++i) {
          d2[i] = mul2(mul2(d2[i], g2[i]), rs);    // dz
          z2[i] = mul2(z2[i], rs);                 // hs
        }
        if (pc == 0) {
          // DZ[g] / HS[g] are free once the MMA that read this group's previous dz tile has retired and the slab stores
          // issued from them have finished reading shared memory (the storing thread waits, the slab's two warps sync)
          ptx::mbar_wait(&dz_empty[g], ((f >> 1) & 1) ^ 1u);
          if (store_thread) ptx::bulk_wait_read<0>();
          named_bar_sync(bar_id, 64);
        }
        const int c16_0 = col0 / 8;
#pragma unroll
        for (int c = 0; c < ZW / 8; ++c) {
          const f32x2 pd[4] = {d2[4 * c], d2[4 * c + 1], d2[4 * c + 2], d2[4 * c + 3]};
          const f32x2 ph[4] = {z2[4 * c], z2[4 * c + 1], z2[4 * c + 2], z2[4 * c + 3]};
          *reinterpret_cast<uint4*>(dzb + sw128(row, c16_0 + c)) = pack8_bf16(pd);
          *reinterpret_cast<uint4*>(hsb + sw128(row, c16_0 + c)) = pack8_bf16(ph);
        }
      }
      ptx::fence_proxy_async();
      named_bar_sync(bar_id, 64);  // the 32 x 64 slab of both tiles is complete
      if (store_thread) {
        ptx::tma_store_2d(&tmDZ, dzb + q * 4096, j * BH, mrow0);   // box {64 hidden, 32 rows}
        ptx::tma_store_2d(&tmHS, hsb + q * 4096, j * BH, mrow0);
        ptx::bulk_commit();
      }
      if (lane == 0) ptx::mbar_arrive(&dz_full[g]);
      if (pending >= 0) {
        write_tile(pending);
        pending = -1;
      }
      if (j == NJ - 1) pending = it;
    }
    if (pending >= 0) write_tile(pending);
    if (lane == 0) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == F_MMA) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

int tmap2d(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer,
           int swizzle) {
  const unsigned long long dims[2] = {(unsigned long long)inner, (unsigned long long)outer};
  const unsigned long long strides[1] = {(unsigned long long)ld * 2ull};
  const unsigned box[2] = {(unsigned)box_inner, (unsigned)box_outer};
  return ogv_make_tmap(tm, ptr, OGV_BF16, 2, dims, strides, box, swizzle);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int C, int BH>
int launch_mlp_fwd(const void* x, long long ldx, const void* w1, const float* b1, const void* w2, const float* b2,
                   const void* residual, long long ldr, const float* row_scale, int rows_per_scale, void* y, long long ldy,
                   long long M, int Hd, int act, cudaStream_t stream) {
  using Cfg = MlpFwdCfg<C, BH>;
  CUtensorMap tmX, tmW1, tmW2, tmY, tmR;
  int rc;
  if ((rc = tmap2d(&tmX, x, C, M, ldx, 64, F_BM, 3))) return rc;
  if ((rc = tmap2d(&tmW1, w1, C, Hd, C, 64, BH, 3))) return rc;
  if ((rc = tmap2d(&tmW2, w2, Hd, C, Hd, 64, C, 3))) return rc;
  if ((rc = tmap2d(&tmY, y, C, M, ldy, 32, 32, 2))) return rc;
  tmR = tmY;
  if (residual && (rc = tmap2d(&tmR, residual, C, M, ldr, 32, 32, 2))) return rc;
  MlpFwdParams p;
  p.M = M; p.Hd = Hd; p.NJ = Hd / BH; p.m_tiles = ogv_ceil_div(M, F_BM); p.act = act; p.has_res = residual ? 1 : 0;
  p.b1 = b1; p.b2 = b2; p.row_scale = row_scale; p.rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(mlp_fwd_kernel<C, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (err != cudaSuccess) {
      ogv_set_error("mlp_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(err));
      return OGV_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = p.m_tiles < ogv_num_sms() ? p.m_tiles : ogv_num_sms();
  mlp_fwd_kernel<C, BH><<<grid, F_THREADS, Cfg::SMEM, stream>>>(tmX, tmW1, tmW2, tmY, tmR, p);
  return ogv_check_launch("mlp_fwd");
}


template <int C>
int launch_mlp_bwd(const void* xn, long long ldx, const void* dy, long long ldg, const void* w1, const void* w2t,
                   const void* w1t, const float* b1, const float* row_scale, int rows_per_scale, void* dz, void* hs,
                   void* dxn, long long lddx, long long M, int Hd, int act, cudaStream_t stream) {
  using Cfg = MlpBwdCfg<C>;
  CUtensorMap tmX, tmG, tmW1, tmW2T, tmW1T, tmDZ, tmHS, tmDX;
  int rc;
  if ((rc = tmap2d(&tmX, xn, C, M, ldx, 64, F_BM, 3))) return rc;
  if ((rc = tmap2d(&tmG, dy, C, M, ldg, 64, F_BM, 3))) return rc;
  if ((rc = tmap2d(&tmW1, w1, C, Hd, C, 64, B_BH, 3))) return rc;
  if ((rc = tmap2d(&tmW2T, w2t, C, Hd, C, 64, B_BH, 3))) return rc;
  if ((rc = tmap2d(&tmW1T, w1t, Hd, C, Hd, 64, C, 3))) return rc;
  if ((rc = tmap2d(&tmDZ, dz, Hd, M, Hd, 64, 32, 3))) return rc;
  if ((rc = tmap2d(&tmHS, hs, Hd, M, Hd, 64, 32, 3))) return rc;
  if ((rc = tmap2d(&tmDX, dxn, C, M, lddx, 32, 32, 2))) return rc;
  MlpBwdParams p;
  p.M = M; p.Hd = Hd; p.NJ = Hd / B_BH; p.m_tiles = ogv_ceil_div(M, F_BM); p.act = act; p.b1 = b1;
  p.row_scale = row_scale; p.rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(mlp_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (err != cudaSuccess) {
      ogv_set_error("mlp_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(err));
      return OGV_ERR_CUDA;
    }
    attr_set = true;
  }
  const int grid = p.m_tiles < ogv_num_sms() ? p.m_tiles : ogv_num_sms();
  mlp_bwd_kernel<C><<<grid, F_THREADS, Cfg::SMEM, stream>>>(tmX, tmG, tmW1, tmW2T, tmW1T, tmDZ, tmHS, tmDX, p);
  return ogv_check_launch("mlp_bwd");
}

// hidden-chunk width of the forward kernel for a given channel count (0: shape not served by the fused kernel)
int fwd_chunk(int C, int Hd) {
  static int bh64 = -1;  // OGV_MLP_BH64=1: 64-wide hidden chunks at C = 64 too (two hidden tiles per group, 6 Z buffers)
  if (bh64 < 0) { const char* e = getenv("OGV_MLP_BH64"); bh64 = (e && e[0] == '1') ? 1 : 0; }
  const int bh = C == 64 ? (bh64 ? 64 : 128) : (C == 128 ? 64 : 0);
  return (bh && Hd % bh == 0 && Hd >= bh && Hd % B_BH == 0) ? bh : 0;
}

}  // namespace

extern "C" int ogv_mlp_fused_supported(int C, int Hd) { return fwd_chunk(C, Hd) ? 1 : 0; }

extern "C" int ogv_mlp_fwd(const void* x, long long ldx, const void* w1, const float* b1, const void* w2, const float* b2,
                           const void* residual, long long ldr, const float* row_scale, int rows_per_scale, void* y,
                           long long ldy, long long M, int C, int Hd, int act, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(x && w1 && b1 && w2 && b2 && y, "mlp_fwd: null argument");
  OGV_REQUIRE(act >= OGV_ACT_NONE && act <= OGV_ACT_RELU, "mlp_fwd: bad activation code %d", act);
  if (!fwd_chunk(C, Hd)) {
    ogv_set_error("mlp_fwd: unsupported shape C=%d hidden=%d (C in {64, 128}, hidden a multiple of the chunk width)", C, Hd);
    return OGV_ERR_UNSUPPORTED;
  }
  OGV_REQUIRE(aligned16(x) && aligned16(w1) && aligned16(w2) && aligned16(y) && aligned16(residual) && aligned16(b1) &&
                  aligned16(b2) && ldx % 8 == 0 && ldy % 8 == 0 && ldr % 8 == 0,
              "mlp_fwd: tensors must be 16-byte aligned with row strides that are multiples of 8 elements");
  OGV_REQUIRE(M <= 0x7fffffffll - F_BM, "mlp_fwd: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 64 && fwd_chunk(C, Hd) == 64)
    return launch_mlp_fwd<64, 64>(x, ldx, w1, b1, w2, b2, residual, ldr, row_scale, rows_per_scale, y, ldy, M, Hd, act, st);
  if (C == 64)
    return launch_mlp_fwd<64, 128>(x, ldx, w1, b1, w2, b2, residual, ldr, row_scale, rows_per_scale, y, ldy, M, Hd, act, st);
  return launch_mlp_fwd<128, 64>(x, ldx, w1, b1, w2, b2, residual, ldr, row_scale, rows_per_scale, y, ldy, M, Hd, act, st);
}

extern "C" int ogv_mlp_bwd(const void* xn, long long ldx, const void* dy, long long ldg, const void* w1, const void* w2t,
                           const void* w1t, const float* b1, const float* row_scale, int rows_per_scale, void* dz, void* hs,
                           void* dxn, long long lddx, long long M, int C, int Hd, int act, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(xn && dy && w1 && w2t && w1t && b1 && dz && hs && dxn, "mlp_bwd: null argument");
  OGV_REQUIRE(act >= OGV_ACT_NONE && act <= OGV_ACT_RELU, "mlp_bwd: bad activation code %d", act);
  if (!fwd_chunk(C, Hd)) {
    ogv_set_error("mlp_bwd: unsupported shape C=%d hidden=%d", C, Hd);
    return OGV_ERR_UNSUPPORTED;
  }
  OGV_REQUIRE(aligned16(xn) && aligned16(dy) && aligned16(w1) && aligned16(w2t) && aligned16(w1t) && aligned16(b1) &&
                  aligned16(dz) && aligned16(hs) && aligned16(dxn) && ldx % 8 == 0 && ldg % 8 == 0 && lddx % 8 == 0,
              "mlp_bwd: tensors must be 16-byte aligned with row strides that are multiples of 8 elements");
  OGV_REQUIRE(M <= 0x7fffffffll - F_BM, "mlp_bwd: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 64)
    return launch_mlp_bwd<64>(xn, ldx, dy, ldg, w1, w2t, w1t, b1, row_scale, rows_per_scale, dz, hs, dxn, lddx, M, Hd, act, st);
  return launch_mlp_bwd<128>(xn, ldx, dy, ldg, w1, w2t, w1t, b1, row_scale, rows_per_scale, dz, hs, dxn, lddx, M, Hd, act, st);
}
