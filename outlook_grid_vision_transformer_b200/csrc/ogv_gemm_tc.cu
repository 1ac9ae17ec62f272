// tcgen05 GEMM engine for sm_100a:  D[M,N] = epi( A[M,K] * B[N,K]^T ), bf16 operands, fp32 accumulate.
//
//   * persistent, warp-specialised CTA of 18 warps, one CTA per SM:
//       warps 0..15  epilogue: tcgen05.ld (lane = output row, warp%4 = TMEM lane quarter, warp/4 =
//                    column group) -> fused epilogue -> swizzled smem staging -> TMA store; the
//                    residual / saved-pre-activation operand of the epilogue arrives by TMA load
//                    into the same staging slot.  (The fused epilogues are ALU/MUFU heavy -- GELU,
//                    GELU' -- so they get 16 warps, 4 per scheduler.)
//       warp 16      TMA producer (cp.async.bulk.tensor 2-D boxes, 128B swizzle, mbarrier complete_tx)
//       warp 17      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, cta_group::1)
//   * smem ring of `stages` x (A 128x64 | B BNx64) bf16 tiles; accumulators double-buffered in TMEM
//     (2 x BN fp32 columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * operands may be K-major (forward / dgrad) or MN-major (wgrad: reduction over the row index
//     of both global tensors) -- only the TMA box and the UMMA descriptors differ;
//   * split-K work items (wgrad) accumulate with fp32 atomics into D (direct, non-TMA epilogue).
#include <stdlib.h>

#include "ogv_gemm.cuh"
#include "ogv_ptx.cuh"
#include "ogv_stage.cuh"
#include "ogv_tma.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = one 128-byte swizzle row
// 16 epilogue warps (4 per scheduler, 96 registers each).  Measured against 8 warps x 168 registers
// (-DOGV_EPI_WARPS=8, no spills): plain-store GEMMs gain ~8 % with 8, the GELU / saved-derivative epilogues
// lose ~10 % (they are ALU-bound and want the warps); the sum over the model's shapes is a wash.
#ifndef OGV_EPI_WARPS
#define OGV_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = OGV_EPI_WARPS;
constexpr int TC_THREADS = (EPI_WARPS + 2) * 32;
constexpr int PRODUCER_WARP = EPI_WARPS;
constexpr int MMA_WARP = EPI_WARPS + 1;
constexpr int CH = 32;                     // epilogue chunk: 32 columns = 64 B of bf16 per row
constexpr int SLOT_BYTES = 32 * CH * 2;    // one warp's 32 x 32 bf16 staging tile (64B-swizzled)
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = (2 * MAX_STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
constexpr int ONES_BYTES = 2048;  // B operand of the row-sum MMA: 16 rows x 64 k of bf16 ones (K-major)

// Patch view: one operand of the GEMM is the (never materialised) im2col matrix of a channels_last bf16 image,
//   cols[(b, oy, ox), (ky*3 + kx)*Cin + c] = x[b, oy*s - 1 + ky, ox*s - 1 + kx, c]        (3x3, pad 1, stride s = 1 | 2)
// A 64-channel slice of one tap for a run of output pixels is ONE 5-D TMA box of x viewed as
//   s = 2: {2*Cin (w-parity, c), W/2, 2 (h-parity), H/2, B}      s = 1: {Cin, W, 1, H, B}
// -- box {64, Wo, 1, rows, images}: whole output rows, so the tile lands in shared memory in the row order of the
// output matrix and with the same 128B-swizzled layout as a 2-D box of the materialised matrix; the taps that fall off
// the image are the descriptor's zero fill (coordinate -1 / past the end).
struct ConvView {
  int which;   // 0: none; 1: the K-major A operand (forward: rows = pixels, k = patch columns);
               // 2: the MN-major B operand (weight gradient: n = patch columns, k = pixels)
  int Cin, stride, Wo, rpi /* Ho*Wo */, B;
};

__device__ __forceinline__ void conv_patch_load(void* dst, const CUtensorMap* tm, uint64_t* bar, const ConvView& cv,
                                                int row0, int col0) {
  const int tap = col0 / cv.Cin, c0 = col0 - tap * cv.Cin;
  const int ky = tap / 3, kx = tap - 3 * ky;
  int b0 = row0 / cv.rpi;
  const int oy0 = (row0 - b0 * cv.rpi) / cv.Wo;
  if (tap >= 9) b0 = cv.B;  // columns past the patch matrix (N tail of a weight-gradient tile): all zero fill
  if (cv.stride == 2)
    ptx::tma_load_5d(dst, tm, bar, (kx != 1 ? cv.Cin : 0) + c0, kx == 0 ? -1 : 0, ky != 1 ? 1 : 0,
                     oy0 + (ky == 0 ? -1 : 0), b0);
  else
    ptx::tma_load_5d(dst, tm, bar, c0, kx - 1, 0, oy0 + ky - 1, b0);
}

struct TcParams {
  int M, N, K;
  ConvView cv;
  int a_mn, b_mn;
  int m_tiles, n_tiles, splits, chunks_per_split, k_chunks;
  int stages;       // smem ring depth
  int slots;        // staging slots per epilogue warp (0: direct epilogue, 2, or 4 with pre_out)
  int epi_load;     // 0 none, 1 residual, 2 dact_src -- operand fetched by TMA into the staging slot
  int vec_ok;       // direct epilogue may use 8-wide vector accesses
  int col_stats;    // staged epilogue also accumulates per-column sum (and sum of squares) of the stored values
  int plain;        // staged epilogue with no bias / activation / operand / scale / statistics: convert and store
  int n_pad;        // n_tiles * BN: extent of the per-CTA column accumulators in shared memory
  float* row_sum;   // optional [M]: += sum_k A(m, k) -- the bias gradient riding with a weight-gradient GEMM
  int rs_col;       // TMEM column of the 16-column row-sum accumulator
  GemmEpi epi;
};

// ---------------------------------------------------------------------------------------------
// direct (non-TMA) epilogue pieces: fp32 output, split-K atomics, unaligned tensors
// ---------------------------------------------------------------------------------------------
template <typename TO>
__device__ __forceinline__ void epi_row16(const GemmEpi& e, int m, int n0, const float* v, int vec_ok) {
  if (vec_ok && n0 + 16 <= e.N) {
    const float rs = e.row_scale ? e.row_scale[m / e.rows_per_scale] : 1.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 8 * h;
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = v[8 * h + i];
      if (e.bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += __ldg(e.bias + n + i);
      }
      if (e.pre_out && e.pre_out_grad) {
        float d[8];
        act_both_n<8, FastAct<TO>::value>(e.act, x, d);
        st8(reinterpret_cast<TO*>(e.pre_out) + (long long)m * e.ld_pre + n, d);
      } else {
        if (e.pre_out) st8(reinterpret_cast<TO*>(e.pre_out) + (long long)m * e.ld_pre + n, x);
        if (e.act != OGV_ACT_NONE) act_apply_n<8, FastAct<TO>::value>(e.act, x);
      }
      if (e.dact_src) {
        float d[8];
        ld8(reinterpret_cast<const TO*>(e.dact_src) + (long long)m * e.ld_dact + n, d);
        act_grad_mul_n<8, FastAct<TO>::value>(e.dact, x, d);
      }
      if (e.row_scale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] *= rs;
      }
      if (e.residual) {
        float r[8];
        ld8(reinterpret_cast<const TO*>(e.residual) + (long long)m * e.ld_res + n, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] += r[i];
      }
      if (e.accumulate) {
        float* d = reinterpret_cast<float*>(e.D) + (long long)m * e.ldd + n;
        ptx::red_add_v4(d, x[0], x[1], x[2], x[3]);
        ptx::red_add_v4(d + 4, x[4], x[5], x[6], x[7]);
      } else {
        st8(reinterpret_cast<TO*>(e.D) + (long long)m * e.ldd + n, x);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < e.N) epi_scalar<TO>(e, m, n0 + j, v[j]);
  }
}

// MT = M sub-tiles (of 128 rows) per work item.  MT = 2 (narrow outputs, BN <= 128): a 128 x 64 output tile keeps only half
// of the epilogue warps busy and pays the per-tile hand-off latency (MMA commit -> epilogue -> TMA store -> accumulator
// free) for very little data; 256 rows per item put all 16 warps to work and halve the number of hand-offs.
template <int BN, typename TO, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmP,
               const __grid_constant__ CUtensorMap tmL, const TcParams p) {
  constexpr int A_BYTES = MT * TC_BM * TC_BK * 2;
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ACC_COLS = MT * BN;        // TMEM columns of one accumulator buffer
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "TMEM budget");
  // the row-sum path exists only in the fp32-output instantiations (weight gradients): the bf16 epilogues, whose
  // register budget is exact, do not carry it
  constexpr bool kRowSum = sizeof(TO) == 4;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment.
  uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones_tile = smem_al;  // only present (and skipped over) when p.row_sum
  uint8_t* smem = smem_al + ((kRowSum && p.row_sum) ? ONES_BYTES : 0);
  uint8_t* staging = smem + p.stages * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + EPI_WARPS * p.slots * SLOT_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* ld_bar = tempty_bar + 2;  // [EPI_WARPS][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + 2 * EPI_WARPS);
  float* colacc = reinterpret_cast<float*>(tmem_slot + 4);  // [2][n_pad] when p.col_stats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == PRODUCER_WARP && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull_bar[s], 1);
      ptx::mbar_init(&tempty_bar[s], EPI_WARPS);
    }
    for (int s = 0; s < 2 * EPI_WARPS; ++s) ptx::mbar_init(&ld_bar[s], 1);
    ptx::fence_barrier_init();
  }
  const uint32_t tmem_cols = (kRowSum && p.row_sum) ? 512u : (uint32_t)TMEM_COLS;  // room for the row-sum accumulator
  if (warp == MMA_WARP) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  if (kRowSum && p.row_sum) {
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(ones_tile)[i] = 0x3f803f80u;
    ptx::fence_proxy_async();
  }
  if (p.col_stats)
    for (int i = threadIdx.x; i < 2 * p.n_pad; i += TC_THREADS) colacc[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = p.m_tiles * p.n_tiles * p.splits;

  if (warp == PRODUCER_WARP) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int nt = w % p.n_tiles;
        const int mt = (w / p.n_tiles) % p.m_tiles;
        const int sp = w / (p.n_tiles * p.m_tiles);
        const int m0 = mt * (MT * TC_BM), n0 = nt * BN;
        const int kc0 = sp * p.chunks_per_split;
        const int kc1 = min(p.k_chunks, kc0 + p.chunks_per_split);
        for (int kc = kc0; kc < kc1; ++kc) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kc * TC_BK;
          if (p.cv.which == 1) {
#pragma unroll
            for (int i = 0; i < MT; ++i)  // 128 output pixels x 64 channels of one tap
              conv_patch_load(sa + i * (TC_BM * TC_BK * 2), &tmA, &full_bar[stage], p.cv, m0 + i * TC_BM, k0);
          } else if (!p.a_mn) {
#pragma unroll
            for (int i = 0; i < MT; ++i)  // box {64 k, 128 rows} per M sub-tile
              ptx::tma_load_2d(sa + i * (TC_BM * TC_BK * 2), &tmA, &full_bar[stage], k0, m0 + i * TC_BM);
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / 64; ++j)  // box {64 mn, 64 k-rows}
              ptx::tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, k0);
          }
          if (p.cv.which == 2) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)  // 64 pixels (k rows) x 64 patch columns
              conv_patch_load(sb + j * 8192, &tmB, &full_bar[stage], p.cv, k0, n0 + 64 * j);
          } else if (!p.b_mn) {
            ptx::tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);  // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, k0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16(TC_BM, BN, p.a_mn, p.b_mn);
      // K-major  : rows of 128 B, 8-row groups 1024 B apart (SBO); +32 B per UMMA_K step.
      // MN-major : 64-element MN blocks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
      //            +2048 B (16 k-rows) per UMMA_K step.
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;
      // row sums of A as one more (N = 16) MMA per K step against a tile of ones: every column of its accumulator
      // holds sum_k A(m, k).  Only the CTAs of the first N tile do it.
      const uint32_t idesc_rs = ptx::umma_idesc_bf16(TC_BM, 16, p.a_mn, 0);
      const uint32_t ones_addr = ptx::smem_u32(ones_tile);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int sp = w / (p.n_tiles * p.m_tiles);
        const int kc0 = sp * p.chunks_per_split;
        const int kc1 = min(p.k_chunks, kc0 + p.chunks_per_split);
        const int as = it & 1;
        const bool do_rs = kRowSum && p.row_sum != nullptr && (w % p.n_tiles) == 0;
        ptx::mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int kc = kc0; kc < kc1; ++kc) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adesc = ptx::umma_smem_desc(sa + k * a_step, a_lbo, 1024u);
            const uint64_t bdesc = ptx::umma_smem_desc(sb + k * b_step, b_lbo, 1024u);
            ptx::umma_f16(d_tmem, adesc, bdesc, idesc, (kc > kc0 || k > 0) ? 1u : 0u);
            if constexpr (MT == 2)  // second 128-row sub-tile: same B, next accumulator (K-major A only)
              ptx::umma_f16(d_tmem + BN, ptx::umma_smem_desc(sa + TC_BM * TC_BK * 2 + k * a_step, a_lbo, 1024u), bdesc,
                            idesc, (kc > kc0 || k > 0) ? 1u : 0u);
            if (do_rs)
              ptx::umma_f16(tmem_base + p.rs_col, adesc, ptx::umma_smem_desc(ones_addr + k * 32u, 16u, 1024u), idesc_rs,
                            (kc > kc0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue warps
      }
    }
  } else {
    // ------------------------------- epilogue -------------------------------
    const int q = warp & 3;    // TMEM lane quarter this warp may read
    const int cg = warp >> 2;  // column group: chunks cg, cg + EPI_WARPS/4, ...
    const GemmEpi& e = p.epi;
    uint8_t* my_slots = staging + warp * p.slots * SLOT_BYTES;
    uint64_t* my_ld = ld_bar + 2 * warp;
    uint32_t ld_phase = 0;  // bit i = parity of my_ld[i]
    int chunk_ctr = 0;
    int it = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int nt = w % p.n_tiles;
      const int mt = (w / p.n_tiles) % p.m_tiles;
      const int m0 = mt * (MT * TC_BM), n0 = nt * BN;
      const int as = it & 1;
      bool waited = false;
#pragma unroll 1
      for (int cc = cg; cc < MT * (BN / CH); cc += EPI_WARPS / 4) {
        const int sub = MT == 1 ? 0 : cc / (BN / CH);   // M sub-tile
        const int c = MT == 1 ? cc : cc - sub * (BN / CH);  // 32-column chunk within the tile
        const int mrow0 = m0 + sub * TC_BM + q * 32;
        const int m = mrow0 + lane;
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC_COLS + sub * BN;
        const int n0c = n0 + c * CH;
        if (n0c >= p.N) {
          if (MT == 1) break;
          continue;
        }
        uint8_t* s0 = nullptr;
        uint8_t* s1 = nullptr;
        int pr = 0;
        if (p.slots) {
          pr = chunk_ctr & 1;
          ++chunk_ctr;
          s0 = my_slots + (p.slots == 4 ? 2 * pr : pr) * SLOT_BYTES;
          s1 = s0 + SLOT_BYTES;
          // the stores issued two chunks ago from this slot pair must have finished reading smem
          if (lane == 0) ptx::bulk_wait_read<1>();
          __syncwarp();
          if (p.epi_load && lane == 0) {
            ptx::mbar_arrive_expect_tx(&my_ld[pr], SLOT_BYTES);
            ptx::tma_load_2d(s0, &tmL, &my_ld[pr], n0c, mrow0);
          }
        }
        if (!waited) {
          ptx::mbar_wait(&tfull_bar[as], (it >> 1) & 1);
          ptx::tc_fence_after();
          waited = true;
        }
        float v[32];
        ptx::tmem_ld32(t0 + c * CH, v);
        if (!p.slots) {
          if (m < p.M) {
            epi_row16<TO>(e, m, n0c, v, p.vec_ok);
            epi_row16<TO>(e, m, n0c + 16, v + 16, p.vec_ok);
          }
          continue;
        }
        // ---- staged epilogue (bf16 tensors, TMA in/out) ----
        if (p.plain) {  // convert and store, nothing else (expand / project / dgrad without a fused derivative)
          stage_write_row(s0, lane, v);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmD, s0, n0c, mrow0);
            ptx::bulk_commit();
          }
          continue;
        }
        // fused epilogues work on element pairs: FADD2 / FMUL2 / FFMA2
        f32x2 v2[16];
        pack_n<16>(v, v2);
        if (e.bias) {
          if (n0c + CH <= p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(e.bias + n0c) + i);
              v2[2 * i] = add2(v2[2 * i], b4.x);
              v2[2 * i + 1] = add2(v2[2 * i + 1], b4.y);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v2[i] = add2(v2[i], pk2(n0c + 2 * i < p.N ? __ldg(e.bias + n0c + 2 * i) : 0.f,
                                      n0c + 2 * i + 1 < p.N ? __ldg(e.bias + n0c + 2 * i + 1) : 0.f));
          }
        }
        if (e.pre_out && e.pre_out_grad) {
          f32x2 d2[16];
          act_both_p<16, FastAct<TO>::value>(e.act, v2, d2);
          stage_write_row(s1, lane, d2);
        } else {
          if (e.pre_out) stage_write_row(s1, lane, v2);
          act_apply_p<16, FastAct<TO>::value>(e.act, v2);
        }
        if (p.epi_load) {
          ptx::mbar_wait(&my_ld[pr], (ld_phase >> pr) & 1u);
          ld_phase ^= 1u << pr;
          f32x2 r2[16];
          stage_read_row(s0, lane, r2);
          if (p.epi_load == 2) {
            act_grad_mul_p<16, FastAct<TO>::value>(e.dact, v2, r2);
            if (e.row_scale) {
              const f32x2 rs = splat2(m < p.M ? e.row_scale[m / e.rows_per_scale] : 0.f);
#pragma unroll
              for (int i = 0; i < 16; ++i) v2[i] = mul2(v2[i], rs);
            }
          } else {
            const f32x2 rs = splat2((e.row_scale && m < p.M) ? e.row_scale[m / e.rows_per_scale] : 1.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) v2[i] = fma2(v2[i], rs, r2[i]);
          }
          __syncwarp();  // every lane has read its residual row before anyone overwrites the slot
        } else if (e.row_scale) {
          const f32x2 rs = splat2(m < p.M ? e.row_scale[m / e.rows_per_scale] : 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) v2[i] = mul2(v2[i], rs);
        }
        stage_write_row(s0, lane, v2);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tmD, s0, n0c, mrow0);
          if (e.pre_out) ptx::tma_store_2d(&tmP, s1, n0c, mrow0);
          ptx::bulk_commit();
        }
        if (p.col_stats) {
          // column statistics of the chunk AS STORED, read back from the staged tile (12 -> 3.5 instructions per
          // element vs. a register butterfly): lanes 0-15 walk the even rows, lanes 16-31 the odd rows, each
          // lane owning the column pair (2j, 2j+1); every warp load is two conflict-free 64-byte rows.
          const int rows_valid = min(32, p.M - mrow0);
          const int j = lane & 15;
          float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f;
#pragma unroll 8
          for (int i = 0; i < 16; ++i) {
            const int row = 2 * i + (lane >> 4);
            if (row < rows_valid) {
              const uint32_t off = row * 64 + (((uint32_t)(j >> 2) ^ ((row >> 1) & 3)) << 4) + (j & 3) * 4;
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(s0 + off));
              sa += f.x; sb += f.y;
              qa = fmaf(f.x, f.x, qa); qb = fmaf(f.y, f.y, qb);
            }
          }
          sa += __shfl_xor_sync(0xffffffffu, sa, 16);
          sb += __shfl_xor_sync(0xffffffffu, sb, 16);
          qa += __shfl_xor_sync(0xffffffffu, qa, 16);
          qb += __shfl_xor_sync(0xffffffffu, qb, 16);
          if (lane < 16) {
            const int ca = n0c + 2 * j;
            if (ca < p.N) atomicAdd(&colacc[ca], sa);
            if (ca + 1 < p.N) atomicAdd(&colacc[ca + 1], sb);
            if (e.col_sumsq) {
              if (ca < p.N) atomicAdd(&colacc[p.n_pad + ca], qa);
              if (ca + 1 < p.N) atomicAdd(&colacc[p.n_pad + ca + 1], qb);
            }
          }
        }
      }
      if (!waited) {  // warps without a chunk in this tile still pace themselves on the accumulator
        ptx::mbar_wait(&tfull_bar[as], (it >> 1) & 1);
        ptx::tc_fence_after();
      }
      if (kRowSum && p.row_sum && nt == 0 && cg == 0) {  // one warp per lane quarter: column 0 of the row-sum accumulator
        float rs16[16];
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + p.rs_col, rs16);
        const int m = m0 + q * 32 + lane;  // (row sums ride with MT = 1 work items only)
        if (m < p.M) atomicAdd(p.row_sum + m, rs16[0]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
    }
    if (p.slots && lane == 0) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (p.col_stats) {  // one global atomic per column per CTA
    for (int i = threadIdx.x; i < p.N; i += TC_THREADS) {
      if (p.epi.col_sum) atomicAdd(p.epi.col_sum + i, colacc[i]);
      if (p.epi.col_sumsq) atomicAdd(p.epi.col_sumsq + i, colacc[p.n_pad + i]);
    }
  }
  if (warp == MMA_WARP) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D bf16 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, outer stride `ld` elements.
int make_tmap(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
              int box_outer, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    ogv_set_error("cuTensorMapEncodeTiled entry point not available");
    return OGV_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ogv_set_error("cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%dx%d ptr=%p", (int)r, inner,
                  outer, ld, box_inner, box_outer, ptr);
    return OGV_ERR_CUDA;
  }
  return OGV_OK;
}

// 0 = K-major (contiguous along k), 1 = MN-major (contiguous along m/n), -1 = not expressible.
int operand_major(long long rs, long long cs) {
  if (cs == 1) return 0;
  if (rs == 1) return 1;
  return -1;
}

bool tma_ok(const void* ptr, long long ld) {
  return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && (ld % 8 == 0);
}

template <int BN, typename TO, int MT>
int launch_tc(const CUtensorMap* tms, TcParams& p, cudaStream_t stream) {
  constexpr int STAGE_BYTES = MT * TC_BM * TC_BK * 2 + BN * TC_BK * 2;
  const int staging = EPI_WARPS * p.slots * SLOT_BYTES;
  const int colbytes = (p.col_stats ? 2 * p.n_pad * 4 : 0) + (p.row_sum ? ONES_BYTES : 0);
  int stages = (SMEM_LIMIT - 1024 - BAR_BYTES - staging - colbytes) / STAGE_BYTES;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) {
    ogv_set_error("gemm_tc: shared memory budget leaves %d pipeline stages", stages);
    return OGV_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const int smem_bytes = stages * STAGE_BYTES + staging + BAR_BYTES + colbytes + 1024;
  static int attr_bytes = 0;
  if (attr_bytes < smem_bytes) {
    cudaError_t err = cudaFuncSetAttribute(gemm_tc_kernel<BN, TO, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           SMEM_LIMIT);
    if (err != cudaSuccess) {
      ogv_set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(err));
      return OGV_ERR_CUDA;
    }
    attr_bytes = SMEM_LIMIT;
  }
  int total = p.m_tiles * p.n_tiles * p.splits;
  int grid = total < ogv_num_sms() ? total : ogv_num_sms();
  // the row-sum accumulator sits behind the two output accumulators; with BN = 256 they fill the TMEM, and it takes
  // the place of the second one -- which is idle exactly when no CTA gets a second work item
  p.rs_col = 2 * BN < 512 ? 2 * BN : BN;
  if (p.row_sum && 2 * BN >= 512 && total > grid) {
    ogv_set_error("gemm_tc: row_sum with a 256-wide tile needs at most one work item per CTA (%d items, %d CTAs)", total, grid);
    return OGV_ERR_UNSUPPORTED;
  }
  gemm_tc_kernel<BN, TO, MT><<<grid, TC_THREADS, smem_bytes, stream>>>(tms[0], tms[1], tms[2], tms[3], tms[4], p);
  return ogv_check_launch("gemm_tc");
}

}  // namespace

bool ogv_gemm_tc_supported(const ogv_gemm_args& a, const char** why) {
  static const char* msg = "";
  auto fail = [&](const char* m) {
    msg = m;
    if (why) *why = msg;
    return false;
  };
  if (a.in_dtype != OGV_BF16) return fail("operands must be bf16");
  if (a.M < 1 || a.N < 1 || a.K < 1) return fail("empty problem");
  int am = operand_major(a.a_rs, a.a_cs), bm = operand_major(a.b_rs, a.b_cs);
  if (am < 0 || bm < 0) return fail("operand has no unit stride");
  long long a_ld = am ? a.a_cs : a.a_rs, b_ld = bm ? a.b_cs : a.b_rs;
  if (a_ld % 8 || b_ld % 8) return fail("leading dimension not a multiple of 8 elements (16 B)");
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15))
    return fail("operand base not 16-byte aligned");
  if (a.col_sum || a.col_sumsq) {
    // fused only in the staged (bf16, TMA-store) epilogue
    if (!a.col_sum) return fail("col_sumsq needs col_sum");
    if (a.out_dtype != OGV_BF16 || a.accumulate) return fail("column statistics need a bf16, non-accumulating output");
    if ((reinterpret_cast<uintptr_t>(a.D) & 15) || (a.ldd % 8)) return fail("column statistics need a TMA-storable output");
    if (a.N > 4096) return fail("column statistics: N too large");
  }
  if (a.split_k > 1 && !a.accumulate) return fail("split_k>1 needs accumulate");
  if (a.row_sum) {
    if (!a.accumulate) return fail("row_sum rides with accumulate (weight-gradient) GEMMs only");
    const int bn = a.N <= 64 ? 64 : (a.N <= 128 ? 128 : 256);
    const long long items = (long long)ogv_ceil_div(a.M, 128) * ogv_ceil_div(a.N, bn) * (a.split_k > 0 ? a.split_k : 1);
    if (bn == 256 && items > ogv_num_sms()) return fail("row_sum with a 256-wide tile needs one work item per CTA");
  }
  if (a.accumulate && a.out_dtype != OGV_F32) return fail("accumulate needs fp32 output");
  return true;
}

// Box geometry of the patch view for tiles of `rows` output pixels: whole output rows of one image, or whole images.
static bool conv_box(const ogv_conv_view& c, int rows, unsigned (&box)[5]) {
  if (c.stride != 1 && c.stride != 2) return false;
  if (c.Cin % 64 != 0 || c.B < 1 || c.H < 1 || c.W < 1) return false;
  if (c.stride == 2 && ((c.H | c.W) & 1)) return false;
  const int Ho = c.H / c.stride, Wo = c.W / c.stride, rpi = Ho * Wo;
  if (Wo > rows || rows % Wo != 0) return false;
  int bh, bb;
  if (rpi >= rows) {
    if (rpi % rows != 0) return false;
    bh = rows / Wo; bb = 1;
  } else {
    if (rows % rpi != 0) return false;
    bh = Ho; bb = rows / rpi;
  }
  box[0] = 64; box[1] = (unsigned)Wo; box[2] = 1; box[3] = (unsigned)bh; box[4] = (unsigned)bb;
  return true;
}

bool ogv_conv_view_supported(const ogv_conv_view& c) {
  unsigned box[5];
  return conv_box(c, 128, box) && conv_box(c, 64, box) && (reinterpret_cast<uintptr_t>(c.x) % 16 == 0);
}

static int make_conv_tmap(CUtensorMap* tm, const ogv_conv_view& c, int rows) {
  unsigned box[5];
  if (!conv_box(c, rows, box)) {
    ogv_set_error("gemm_tc: patch view not expressible (B=%d H=%d W=%d Cin=%d stride=%d, %d-row tiles)", c.B, c.H, c.W,
                  c.Cin, c.stride, rows);
    return OGV_ERR_UNSUPPORTED;
  }
  const unsigned long long C = (unsigned long long)c.Cin, W = (unsigned long long)c.W, H = (unsigned long long)c.H;
  unsigned long long dims[5], str[4];
  if (c.stride == 2) {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = (unsigned long long)c.B;
    str[0] = 2 * C * 2; str[1] = W * C * 2; str[2] = 2 * W * C * 2; str[3] = H * W * C * 2;
  } else {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = (unsigned long long)c.B;
    str[0] = C * 2; str[1] = W * C * 2; str[2] = W * C * 2; str[3] = H * W * C * 2;
  }
  return ogv_make_tmap(tm, c.x, OGV_BF16, 5, dims, str, box, 3);
}

static int gemm_tc_run(const ogv_gemm_args& a, const ogv_conv_view* conv, cudaStream_t stream);

int ogv_gemm_tc(const ogv_gemm_args& a, cudaStream_t stream) { return gemm_tc_run(a, nullptr, stream); }

int ogv_gemm_tc_conv(const ogv_gemm_args& a, const ogv_conv_view& conv, cudaStream_t stream) {
  return gemm_tc_run(a, &conv, stream);
}

static int gemm_tc_run(const ogv_gemm_args& a, const ogv_conv_view* conv, cudaStream_t stream) {
  const char* why = "";
  if (!ogv_gemm_tc_supported(a, &why)) {
    ogv_set_error("gemm_tc: unsupported problem: %s", why);
    return OGV_ERR_UNSUPPORTED;
  }
  if (a.out_dtype != OGV_BF16 && a.out_dtype != OGV_F32) {
    ogv_set_error("gemm_tc: bad out dtype %d", a.out_dtype);
    return OGV_ERR_ARG;
  }
  TcParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.a_mn = operand_major(a.a_rs, a.a_cs);
  p.b_mn = operand_major(a.b_rs, a.b_cs);
  p.cv.which = 0;
  if (conv && conv->which) {
    if ((conv->which == 1 && p.a_mn != 0) || (conv->which == 2 && p.b_mn != 1) || (conv->which != 1 && conv->which != 2)) {
      ogv_set_error("gemm_tc: patch view on an operand of the wrong major-ness");
      return OGV_ERR_ARG;
    }
    p.cv.which = conv->which; p.cv.Cin = conv->Cin; p.cv.stride = conv->stride; p.cv.B = conv->B;
    p.cv.Wo = conv->W / conv->stride; p.cv.rpi = (conv->H / conv->stride) * p.cv.Wo;
  }
  const int BN = a.N <= 64 ? 64 : (a.N <= 128 ? 128 : 256);
  // two 128-row sub-tiles per work item for narrow bf16 outputs with enough rows to keep every SM busy
  static int mt_mode = -1;  // OGV_GEMM_MT=1 forces single sub-tiles (A/B measurements)
  if (mt_mode < 0) { const char* e = getenv("OGV_GEMM_MT"); mt_mode = e ? atoi(e) : 2; }
  p.m_tiles = ogv_ceil_div(a.M, TC_BM);
  p.n_tiles = ogv_ceil_div(a.N, BN);
  p.k_chunks = ogv_ceil_div(a.K, TC_BK);
  int split = a.split_k > 0 ? a.split_k : 1;
  if (split > p.k_chunks) split = p.k_chunks;
  p.chunks_per_split = ogv_ceil_div(p.k_chunks, split);
  p.splits = ogv_ceil_div(p.k_chunks, p.chunks_per_split);
  p.epi = make_epi(a);
  p.stages = 0;
  p.row_sum = a.row_sum;
  p.rs_col = 0;

  // staged (TMA) epilogue: bf16 tensors, 16-byte aligned rows, at most one epilogue input operand
  const bool obf = a.out_dtype == OGV_BF16;
  const bool staged = obf && !a.accumulate && tma_ok(a.D, a.ldd) && (!a.pre_out || tma_ok(a.pre_out, a.ld_pre)) &&
                      (!a.residual || tma_ok(a.residual, a.ld_res)) && (!a.dact_src || tma_ok(a.dact_src, a.ld_dact)) &&
                      !(a.residual && a.dact_src);
  // the implicit convolution is bound by the L2 -> shared-memory fill (a weight tile per 128-pixel tile): two pixel
  // sub-tiles per weight tile also at BN = 128 (OGV_CONV_MT=1 keeps single sub-tiles, for A/B measurements)
  static int conv_mt = -1;
  if (conv_mt < 0) { const char* e = getenv("OGV_CONV_MT"); conv_mt = e ? atoi(e) : 2; }
  const int mt_bn_max = (mt_mode >= 3 || (p.cv.which == 1 && conv_mt >= 2)) ? 128 : 64;
  const int MT = (mt_mode >= 2 && BN <= mt_bn_max && staged && p.a_mn == 0 && !a.row_sum && a.M >= 2 * 256 * ogv_num_sms()) ? 2 : 1;
  p.m_tiles = ogv_ceil_div(a.M, TC_BM * MT);
  p.slots = staged ? (a.pre_out ? 4 : 2) : 0;
  p.col_stats = (a.col_sum != nullptr) ? 1 : 0;
  p.n_pad = p.n_tiles * BN;
  if (p.col_stats && !staged) {
    ogv_set_error("gemm_tc: column statistics requested on a problem without a staged epilogue");
    return OGV_ERR_UNSUPPORTED;
  }
  p.epi_load = staged ? (a.residual ? 1 : (a.dact_src ? 2 : 0)) : 0;
  p.plain = staged && !p.col_stats && !p.epi_load && !a.pre_out && !a.bias && !a.row_scale && a.act == OGV_ACT_NONE;
  const int esz = obf ? 2 : 4;
  auto ok = [&](const void* ptr, long long ld) {
    return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && ((ld * esz) % 16 == 0));
  };
  // accumulate: D is fp32 whatever TO is, and red.global.v4.f32 needs 16-byte aligned addresses (D base and ldd*4)
  const bool acc_ok = (reinterpret_cast<uintptr_t>(a.D) % 16 == 0) && (a.ldd % 4 == 0);
  p.vec_ok = (a.accumulate ? acc_ok : ok(a.D, a.ldd)) && ok(a.pre_out, a.ld_pre) && ok(a.dact_src, a.ld_dact) &&
             ok(a.residual, a.ld_res);

  CUtensorMap tms[5];
  int rc;
  if (p.cv.which == 1) rc = make_conv_tmap(&tms[0], *conv, TC_BM);
  else if (!p.a_mn) rc = make_tmap(&tms[0], a.A, a.K, a.M, a.a_rs, 64, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = make_tmap(&tms[0], a.A, a.M, a.K, a.a_cs, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (p.cv.which == 2) rc = make_conv_tmap(&tms[1], *conv, 64);
  else if (!p.b_mn) rc = make_tmap(&tms[1], a.B, a.K, a.N, a.b_rs, 64, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = make_tmap(&tms[1], a.B, a.N, a.K, a.b_cs, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tms[2] = tms[0]; tms[3] = tms[0]; tms[4] = tms[0];  // placeholders, never dereferenced when unused
  if (staged) {
    if ((rc = make_tmap(&tms[2], a.D, a.N, a.M, a.ldd, CH, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (a.pre_out && (rc = make_tmap(&tms[3], a.pre_out, a.N, a.M, a.ld_pre, CH, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (a.residual && (rc = make_tmap(&tms[4], a.residual, a.N, a.M, a.ld_res, CH, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (a.dact_src && (rc = make_tmap(&tms[4], a.dact_src, a.N, a.M, a.ld_dact, CH, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }

  switch (BN) {
    case 64:
      if (MT == 2) return launch_tc<64, bf16, 2>(tms, p, stream);
      return obf ? launch_tc<64, bf16, 1>(tms, p, stream) : launch_tc<64, float, 1>(tms, p, stream);
    case 128:
      if (MT == 2) return launch_tc<128, bf16, 2>(tms, p, stream);
      return obf ? launch_tc<128, bf16, 1>(tms, p, stream) : launch_tc<128, float, 1>(tms, p, stream);
    default: return obf ? launch_tc<256, bf16, 1>(tms, p, stream) : launch_tc<256, float, 1>(tms, p, stream);
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 / pad 1 / stride 1|2 convolution over channels_last bf16 rows as an implicit GEMM (include/ogv.h)
// ---------------------------------------------------------------------------------------------
extern "C" int ogv_conv3x3_supported(int B, int H, int W, int Cin, int Co, int stride) {
  ogv_conv_view c{1, reinterpret_cast<const void*>(uintptr_t(16)), B, H, W, Cin, stride};
  return (Co > 0 && Co % 8 == 0 && ogv_conv_view_supported(c)) ? 1 : 0;
}

extern "C" int ogv_conv3x3_fwd(const void* x, const void* w2, void* y, float* col_sum, float* col_sumsq, int B, int H,
                               int W, int Cin, int Co, int stride, void* stream) {
  OGV_REQUIRE(x && w2 && y, "conv3x3_fwd: null pointer");
  ogv_conv_view c{1, x, B, H, W, Cin, stride};
  if (!(Co > 0 && Co % 8 == 0 && ogv_conv_view_supported(c))) {
    ogv_set_error("conv3x3_fwd: unsupported geometry B=%d H=%d W=%d Cin=%d Co=%d stride=%d", B, H, W, Cin, Co, stride);
    return OGV_ERR_UNSUPPORTED;
  }
  ogv_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.A = x; a.a_rs = 9LL * Cin; a.a_cs = 1;
  a.B = w2; a.b_rs = 9LL * Cin; a.b_cs = 1;
  a.D = y; a.ldd = Co;
  a.M = B * (H / stride) * (W / stride); a.N = Co; a.K = 9 * Cin;
  a.in_dtype = OGV_BF16; a.out_dtype = OGV_BF16;
  a.rows_per_scale = 1; a.split_k = 1;
  a.col_sum = col_sum; a.col_sumsq = col_sumsq;
  return ogv_gemm_tc_conv(a, c, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ogv_conv3x3_wgrad(const void* x, const void* dy, float* dw2, int B, int H, int W, int Cin, int Co,
                                 int stride, void* stream) {
  OGV_REQUIRE(x && dy && dw2, "conv3x3_wgrad: null pointer");
  ogv_conv_view c{2, x, B, H, W, Cin, stride};
  if (!(Co > 0 && Co % 8 == 0 && ogv_conv_view_supported(c))) {
    ogv_set_error("conv3x3_wgrad: unsupported geometry B=%d H=%d W=%d Cin=%d Co=%d stride=%d", B, H, W, Cin, Co, stride);
    return OGV_ERR_UNSUPPORTED;
  }
  const int Mo = B * (H / stride) * (W / stride);
  ogv_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.A = dy; a.a_rs = 1; a.a_cs = Co;          // A(m = co, k = pixel) = dy[pixel, co]
  a.B = x; a.b_rs = 1; a.b_cs = 9LL * Cin;    // B(n = patch column, k = pixel): the patch view
  a.D = dw2; a.ldd = 9LL * Cin;
  a.M = Co; a.N = 9 * Cin; a.K = Mo;
  a.in_dtype = OGV_BF16; a.out_dtype = OGV_F32;
  a.rows_per_scale = 1; a.accumulate = 1;
  // one split-K work item per CTA at most (every item ends with a full tile of fp32 reductions)
  const int tiles = ogv_ceil_div(Co, 128) * ogv_ceil_div(9 * Cin, 256);
  int split = ogv_num_sms() / tiles;
  if (split > ogv_ceil_div(Mo, 256)) split = ogv_ceil_div(Mo, 256);
  a.split_k = split < 1 ? 1 : split;
  return ogv_gemm_tc_conv(a, c, reinterpret_cast<cudaStream_t>(stream));
}
