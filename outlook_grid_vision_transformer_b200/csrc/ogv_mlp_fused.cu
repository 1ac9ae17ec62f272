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
constexpr int F_NW = 4;                 // weight-chunk ring depth
constexpr int F_SLOT = 32 * 32 * 2;     // one warp's 32 x 32 bf16 staging tile (SWIZZLE_64B)

template <int C, int BH>
struct MlpFwdCfg {
  static_assert(C % 64 == 0 && BH % 64 == 0, "tiles are cut in 64-column (128-byte) sub-tiles");
  static constexpr int KC = C / 64;
  static constexpr int KH = BH / 64;
  static constexpr int XA_BYTES = F_BM * C * 2;
  static constexpr int H_BYTES = F_BM * BH * 2;
  static constexpr int W_BYTES = BH * C * 2;
  static constexpr int OUT_WARPS = 4 * (C / 32);          // (lane quarter, 32-column chunk) pairs of the output tile
  static constexpr int STAGE_BYTES = OUT_WARPS * F_SLOT;
  static constexpr int NBAR = 2 + 2 + 2 * F_NW + 2 + 2 + 2 + 2 + 2 + 2 + F_EPI_WARPS;
  static constexpr int SMEM = 2 * XA_BYTES + 2 * H_BYTES + F_NW * W_BYTES + STAGE_BYTES + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 512;                   // 2*BH (Z) + 2*C (Y) = 384 -> next power of two
  static_assert(2 * BH + 2 * C <= 512, "TMEM budget");
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

template <int C, int BH>
__global__ void __launch_bounds__(F_THREADS, 1)
mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmR, const MlpFwdParams p) {
  using Cfg = MlpFwdCfg<C, BH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xa = smem;                                   // [2][XA_BYTES]
  uint8_t* hbuf = xa + 2 * Cfg::XA_BYTES;               // [2][H_BYTES]
  uint8_t* wring = hbuf + 2 * Cfg::H_BYTES;             // [F_NW][W_BYTES]
  uint8_t* staging = wring + F_NW * Cfg::W_BYTES;       // [OUT_WARPS][F_SLOT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGE_BYTES);
  uint64_t* xa_full = bars;
  uint64_t* xa_empty = xa_full + 2;
  uint64_t* w_full = xa_empty + 2;
  uint64_t* w_empty = w_full + F_NW;
  uint64_t* z_full = w_empty + F_NW;
  uint64_t* z_empty = z_full + 2;
  uint64_t* h_full = z_empty + 2;
  uint64_t* h_empty = h_full + 2;
  uint64_t* y_full = h_empty + 2;
  uint64_t* y_empty = y_full + 2;
  uint64_t* ld_bar = y_empty + 2;                       // [F_EPI_WARPS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + F_EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == F_PRODUCER && lane == 0) {
    ptx::tma_prefetch_desc(&tmX);
    ptx::tma_prefetch_desc(&tmW1);
    ptx::tma_prefetch_desc(&tmW2);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&xa_full[s], 1);
      ptx::mbar_init(&xa_empty[s], 1);
      ptx::mbar_init(&z_full[s], 1);
      ptx::mbar_init(&z_empty[s], F_EPI_WARPS);
      ptx::mbar_init(&h_full[s], F_EPI_WARPS);
      ptx::mbar_init(&h_empty[s], 1);
      ptx::mbar_init(&y_full[s], 1);
      ptx::mbar_init(&y_empty[s], Cfg::OUT_WARPS);
    }
    for (int s = 0; s < F_NW; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < F_EPI_WARPS; ++s) ptx::mbar_init(&ld_bar[s], 1);
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
  const uint32_t tmem_z = tmem_base;             // Z[zb] at column zb * BH
  const uint32_t tmem_y = tmem_base + 2 * BH;    // Y[yb] at column 2*BH + yb * C
  const int NJ = p.NJ;

  if (warp == F_PRODUCER) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t wc = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int xb = it & 1;
        ptx::mbar_wait(&xa_empty[xb], ((it >> 1) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&xa_full[xb], Cfg::XA_BYTES);
#pragma unroll
        for (int ks = 0; ks < Cfg::KC; ++ks)  // box {64 k, 128 rows}
          ptx::tma_load_2d(xa + xb * Cfg::XA_BYTES + ks * (F_BM * 128), &tmX, &xa_full[xb], ks * 64, tile * F_BM);
        for (int step = 0; step <= NJ; ++step) {
          if (step < NJ) {  // W1 rows [step*BH, +BH): box {64 k, BH rows} per 64-column slice of C
            const int ws = wc % F_NW;
            ptx::mbar_wait(&w_empty[ws], ((wc / F_NW) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::W_BYTES);
#pragma unroll
            for (int ks = 0; ks < Cfg::KC; ++ks)
              ptx::tma_load_2d(wring + ws * Cfg::W_BYTES + ks * (BH * 128), &tmW1, &w_full[ws], ks * 64, step * BH);
            ++wc;
          }
          if (step >= 1) {  // W2[:, (step-1)*BH ..): box {64 hidden, C rows} per 64-column slice of the chunk
            const int ws = wc % F_NW;
            ptx::mbar_wait(&w_empty[ws], ((wc / F_NW) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::W_BYTES);
#pragma unroll
            for (int hs = 0; hs < Cfg::KH; ++hs)
              ptx::tma_load_2d(wring + ws * Cfg::W_BYTES + hs * (C * 128), &tmW2, &w_full[ws], (step - 1) * BH + hs * 64, 0);
            ++wc;
          }
        }
      }
    }
  } else if (warp == F_MMA) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc1 = ptx::umma_idesc_bf16(F_BM, BH, 0, 0);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(F_BM, C, 0, 0);
      uint32_t wc = 0, zc = 0, hc = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int xb = it & 1, yb = it & 1;
        ptx::mbar_wait(&y_empty[yb], ((it >> 1) & 1) ^ 1u);
        ptx::mbar_wait(&xa_full[xb], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t xa_addr = ptx::smem_u32(xa + xb * Cfg::XA_BYTES);
        for (int step = 0; step <= NJ; ++step) {
          if (step < NJ) {  // fc1 of chunk `step`
            const int zb = zc & 1;
            ptx::mbar_wait(&z_empty[zb], ((zc >> 1) & 1) ^ 1u);
            const int ws = wc % F_NW;
            ptx::mbar_wait(&w_full[ws], (wc / F_NW) & 1);
            ptx::tc_fence_after();
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
            ++zc;
            if (step == NJ - 1) ptx::umma_commit(&xa_empty[xb]);
          }
          if (step >= 1) {  // fc2 of chunk `step - 1`
            const int hb = hc & 1;
            ptx::mbar_wait(&h_full[hb], (hc >> 1) & 1);
            const int ws = wc % F_NW;
            ptx::mbar_wait(&w_full[ws], (wc / F_NW) & 1);
            ptx::tc_fence_after();
            const uint32_t ha = ptx::smem_u32(hbuf + hb * Cfg::H_BYTES);
            const uint32_t wb = ptx::smem_u32(wring + ws * Cfg::W_BYTES);
#pragma unroll
            for (int hs = 0; hs < Cfg::KH; ++hs)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = ptx::umma_smem_desc(ha + hs * (F_BM * 128) + kk * 32, 16u, 1024u);
                const uint64_t bd = ptx::umma_smem_desc(wb + hs * (C * 128) + kk * 32, 16u, 1024u);
                ptx::umma_f16(tmem_y + yb * C, ad, bd, idesc2, (step > 1 || hs > 0 || kk > 0) ? 1u : 0u);
              }
            ptx::umma_commit(&w_empty[ws]);
            ++wc;
            ptx::umma_commit(&h_empty[hb]);
            ++hc;
            if (step == NJ) ptx::umma_commit(&y_full[yb]);
          }
        }
      }
    }
  } else {
    // ------------------------------- epilogue warps -------------------------------
    constexpr int ZW = BH / 4;  // hidden columns of a chunk this warp converts (32 or 16)
    const int q = warp & 3;
    const int cg = warp >> 2;
    const bool out_warp = cg < C / 32;
    uint8_t* slot = staging + (q * (C / 32) + (out_warp ? cg : 0)) * F_SLOT;
    uint64_t* my_ld = &ld_bar[warp];
    uint32_t zc = 0, hc = 0, ldc = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const int yb = it & 1;
      const int mrow0 = tile * F_BM + q * 32;
      const long long m = (long long)mrow0 + lane;
      if (out_warp && p.has_res) {  // the residual tile arrives while the chunks are being processed
        if (lane == 0) {
          ptx::bulk_wait_read<0>();  // the previous tile's store has finished reading the slot
          ptx::mbar_arrive_expect_tx(my_ld, F_SLOT);
          ptx::tma_load_2d(slot, &tmR, my_ld, cg * 32, mrow0);
        }
      }
      for (int j = 0; j < NJ; ++j) {
        const int zb = zc & 1;
        ptx::mbar_wait(&z_full[zb], (zc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_z + (static_cast<uint32_t>(q * 32) << 16) + zb * BH + cg * ZW;
        float v[ZW];
        if constexpr (ZW == 32) ptx::tmem_ld32(taddr, v);
        else ptx::tmem_ld16(taddr, v);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&z_empty[zb]);
        ++zc;
        f32x2 v2[ZW / 2];
        pack_n<ZW / 2>(v, v2);
        const float* b1 = p.b1 + j * BH + cg * ZW;
#pragma unroll
        for (int i = 0; i < ZW / 4; ++i) {
          const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(b1) + i);
          v2[2 * i] = add2(v2[2 * i], b4.x);
          v2[2 * i + 1] = add2(v2[2 * i + 1], b4.y);
        }
        act_apply_p<ZW / 2, true>(p.act, v2);
        const int hb = hc & 1;
        ptx::mbar_wait(&h_empty[hb], ((hc >> 1) & 1) ^ 1u);
        // row `q*32 + lane` of the chunk, columns [cg*ZW, +ZW): 16-byte pieces of a 128B-swizzled K-major tile
        uint8_t* ht = hbuf + hb * Cfg::H_BYTES + ((cg * ZW) / 64) * (F_BM * 128);
        const int c16_0 = ((cg * ZW) % 64) / 8;
        const int row = q * 32 + lane;
#pragma unroll
        for (int c = 0; c < ZW / 8; ++c) {
          const f32x2 piece[4] = {v2[4 * c], v2[4 * c + 1], v2[4 * c + 2], v2[4 * c + 3]};
          *reinterpret_cast<uint4*>(ht + sw128(row, c16_0 + c)) = pack8_bf16(piece);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&h_full[hb]);
        ++hc;
      }
      if (out_warp) {
        ptx::mbar_wait(&y_full[yb], (it >> 1) & 1);
        ptx::tc_fence_after();
        float v[32];
        ptx::tmem_ld32(tmem_y + (static_cast<uint32_t>(q * 32) << 16) + yb * C + cg * 32, v);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&y_empty[yb]);
        f32x2 v2[16];
        pack_n<16>(v, v2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const ulonglong2 b4 = __ldg(reinterpret_cast<const ulonglong2*>(p.b2 + cg * 32) + i);
          v2[2 * i] = add2(v2[2 * i], b4.x);
          v2[2 * i + 1] = add2(v2[2 * i + 1], b4.y);
        }
        const float rsf = (p.row_scale && m < p.M) ? p.row_scale[m / p.rows_per_scale] : 1.f;
        const f32x2 rs = splat2(rsf);
        if (p.has_res) {
          ptx::mbar_wait(my_ld, ldc & 1u);
          ++ldc;
          f32x2 r2[16];
          stage_read_row(slot, lane, r2);
#pragma unroll
          for (int i = 0; i < 16; ++i) v2[i] = fma2(v2[i], rs, r2[i]);
          __syncwarp();  // every lane has read its residual row before anyone overwrites the slot
        } else {
          if (p.row_scale) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v2[i] = mul2(v2[i], rs);
          }
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
        }
        stage_write_row(slot, lane, v2);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tmY, slot, cg * 32, mrow0);
          ptx::bulk_commit();
        }
      }
    }
    if (out_warp && lane == 0) ptx::bulk_wait<0>();
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

// hidden-chunk width of the forward kernel for a given channel count (0: shape not served by the fused kernel)
int fwd_chunk(int C, int Hd) {
  const int bh = C == 64 ? 128 : (C == 128 ? 64 : 0);
  return (bh && Hd % bh == 0 && Hd >= bh) ? bh : 0;
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
  if (C == 64)
    return launch_mlp_fwd<64, 128>(x, ldx, w1, b1, w2, b2, residual, ldr, row_scale, rows_per_scale, y, ldy, M, Hd, act, st);
  return launch_mlp_fwd<128, 64>(x, ldx, w1, b1, w2, b2, residual, ldr, row_scale, rows_per_scale, y, ldy, M, Hd, act, st);
}
