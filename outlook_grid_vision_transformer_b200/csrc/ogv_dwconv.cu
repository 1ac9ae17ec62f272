// Depthwise 3x3 convolution of MBConv (mbc_conv.py:73-78) with the BatchNorm-affine + activation of
// the expand stage fused on load and the batch statistics of the output fused on store, and its
// backward (input gradient with act', filter gradient, BN1 reductions).  Math: SURVEY Appendix A.3/A.4.
//
// Structure (both directions): a persistent CTA owns one CC-channel chunk and walks spatial tiles.
// The halo tile of every step is staged by TMA (cp.async.bulk.tensor.4d over [C, W, H, B]; image
// borders and ragged edges are the descriptor's zero fill, so there is no address arithmetic or
// predication on the load side) into a raw slot; a first pass turns it into an fp32 tile (forward:
// BN1 + activation evaluated ONCE per element -- it is MUFU-bound and the 3x3 stencil would
// otherwise evaluate it 9 times; backward: plain widening), which frees the raw slot for the TMA
// load of the next tile while the fp32 stencil (4 channels x R rows per thread, no conversions, no
// index decoding) runs.  Per-channel statistics and filter-gradient partial sums stay in registers
// until one flush per CTA.
#include <stdlib.h>

#include "ogv_common.cuh"
#include "ogv_ptx.cuh"
#include "ogv_tma.cuh"
#include "../../include/ogv.h"

namespace {

constexpr int DW_THREADS = 256;
constexpr int DW_BUF_POS = 352;  // halo positions per slot (x CC channels): 10 x 34 for 8 x 32 tiles, 3 CTAs/SM forward

struct DwGeom {
  int B, H, W, Cm;
  int TH, TW, NI, TH2, TW2;  // tile rows / cols / images, halo extents
  int tiles_h, tiles_w;
  int ntiles, nchunks, nworkers;
  FastDiv d_tw2, d_th2, d_tw;
  FastDiv d_tiles_w, d_tiles_h;   // tile decode (valid while ntiles < 65536: `small`)
  FastDiv d_ns2, d_ns4;           // strips of 2 / 4 output rows per tile (backward / tiled forward)
  int small;
  int dbg;  // OGV_DW_DBG bit mask (bottleneck hunting only): 1 skip activation math, 2 skip stencil, 4 skip stores
};

// tiles: full-width rows of up to 32 columns; several whole images per tile when an image is small.
int dw_make_geom(int B, int H, int W, int Cm, int CC, int ctas_per_sm, DwGeom* g) {
  g->B = B; g->H = H; g->W = W; g->Cm = Cm;
  g->TW = W < 32 ? W : 32;
  g->TW2 = g->TW + 2;
  int max_th = DW_BUF_POS / g->TW2 - 2;
  if (max_th < 1) return -1;
  g->TH = H < max_th ? H : max_th;
  if (g->TH >= 4 && g->TH < H) g->TH &= ~3;
  g->TH2 = g->TH + 2;
  g->tiles_h = (H + g->TH - 1) / g->TH;
  g->tiles_w = (W + g->TW - 1) / g->TW;
  g->NI = 1;
  if (g->tiles_h == 1 && g->tiles_w == 1) {
    g->NI = DW_BUF_POS / (g->TH2 * g->TW2);
    if (g->NI > B) g->NI = B;
    if (g->NI > 256) g->NI = 256;
    if (g->NI < 1) g->NI = 1;
  }
  long long groups = (B + g->NI - 1) / g->NI;
  long long nt = groups * g->tiles_h * g->tiles_w;
  if (nt > 0x7fffffff) return -1;
  g->ntiles = (int)nt;
  g->nchunks = (Cm + CC - 1) / CC;
  // workers per channel chunk: nchunks * nworkers CTAs must FIT the resident slots (one wave).  Rounding up put 304 CTAs
  // on 296 slots at Cm = 512 (16 chunks x 19 workers): eight CTAs ran alone in a second wave and the kernel took twice
  // its time (stage 1: 626 us for half the elements of stage 0's 754 us).
  long long want = ((long long)ogv_num_sms() * ctas_per_sm) / g->nchunks;
  if (want > nt) want = nt;
  if (want < 1) want = 1;
  g->nworkers = (int)want;
  g->d_tw2 = make_fastdiv(g->TW2);
  g->d_th2 = make_fastdiv(g->TH2);
  g->d_tw = make_fastdiv(g->TW);
  g->d_tiles_w = make_fastdiv(g->tiles_w);
  g->d_tiles_h = make_fastdiv(g->tiles_h);
  g->d_ns2 = make_fastdiv((g->TH + 1) / 2);
  g->d_ns4 = make_fastdiv((g->TH + 3) / 4);
  g->small = g->ntiles < 65536;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("OGV_DW_DBG"); dbg = e ? atoi(e) : 0; }
    g->dbg = dbg;
  }
  return 0;
}

__device__ __forceinline__ void dw_decode_tile(const DwGeom& g, int t, int& b0, int& h0, int& w0) {
  // t = (grp * tiles_h + th) * tiles_w + tw ; tile counts per image are tiny, grp may be large
  // (runtime-divisor integer divisions cost ~5 % of the backward kernel's instructions: multiply-high instead)
  const int q = g.small ? fdiv(t, g.d_tiles_w) : t / g.tiles_w;
  const int tw_i = t - q * g.tiles_w;
  const int grp = g.small ? fdiv(q, g.d_tiles_h) : q / g.tiles_h;
  const int th_i = q - grp * g.tiles_h;
  b0 = grp * g.NI;
  h0 = th_i * g.TH;
  w0 = tw_i * g.TW;
}

template <typename T>
int dw_tmap(CUtensorMap* tm, const void* ptr, int dtype, const DwGeom& g, int CC, int halo) {
  const unsigned long long es = sizeof(T);
  unsigned long long dims[4] = {(unsigned long long)g.Cm, (unsigned long long)g.W, (unsigned long long)g.H,
                                (unsigned long long)g.B};
  unsigned long long str[3] = {g.Cm * es, (unsigned long long)g.W * g.Cm * es, (unsigned long long)g.H * g.W * g.Cm * es};
  unsigned box[4] = {(unsigned)CC, (unsigned)(g.TW + 2 * halo), (unsigned)(g.TH + 2 * halo), (unsigned)g.NI};
  return ogv_make_tmap(tm, ptr, dtype, 4, dims, str, box, 0);
}

// thread -> stencil work mapping.  "Regular" tiles (TW * NTV divides the CTA): a thread keeps one
// (column x, channel group tv) for the whole kernel and only walks (image, strip) pairs, so the
// stencil loop carries no index decoding at all.
struct DwItems {
  int regular, lanes_per_row, rows_par;
};
template <int NTV>
__host__ __device__ inline DwItems dw_items(int TW) {
  DwItems m;
  m.lanes_per_row = TW * NTV;
  m.regular = (m.lanes_per_row <= DW_THREADS) && (DW_THREADS % m.lanes_per_row == 0);
  m.rows_par = m.regular ? DW_THREADS / m.lanes_per_row : 1;
  return m;
}

// ------------------------------------------------------------------------------------------------
// forward: d_pre[p,c] = sum_t w[c,t] * act(scale1*e_pre + shift1)[p + d_t, c]; stats of d_pre
// smem: raw TMA slot (T) | fp32 activated tile | weights | barrier.  Per tile:
//   wait raw -> BN1+act raw -> fp32 tile (zeros outside the image) -> sync -> TMA(next tile) -> stencil -> sync
// ------------------------------------------------------------------------------------------------
template <typename T, int CC, int ACT>
__global__ void __launch_bounds__(DW_THREADS, 3)
dwconv_fwd_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ scale1,
                  const float* __restrict__ shift1, const float* __restrict__ wgt, T* __restrict__ d_pre,
                  float* __restrict__ sum2, float* __restrict__ sumsq2, const DwGeom g) {
  constexpr int R = 4;                    // output rows per thread item
  constexpr int NTV = CC / 4;             // 4-channel groups per position (stencil pass)
  constexpr int AV = 16 / sizeof(T);      // channels per 16-byte vector (activation pass)
  constexpr int NAV = CC / AV;
  constexpr int BUF_ELEMS = DW_BUF_POS * CC;
  extern __shared__ __align__(128) uint8_t dsm[];
  T* const raw = reinterpret_cast<T*>(dsm);
  float* const tile = reinterpret_cast<float*>(dsm + BUF_ELEMS * sizeof(T));
  uint64_t* const bar = reinterpret_cast<uint64_t*>(dsm + BUF_ELEMS * (sizeof(T) + sizeof(float)));
  __shared__ float s_sum[CC], s_sq[CC];
  __shared__ __align__(16) float s_w[9][CC];

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % g.nchunks;
  const int worker = blockIdx.x / g.nchunks;
  const int c0 = chunk * CC;
  const int tv = tid % NTV;
  const int c = c0 + tv * 4;
  const bool cvalid = c < g.Cm;
  const int av = tid % NAV;
  const int ca = c0 + av * AV;
  const bool cavalid = ca < g.Cm;

  if (tid < CC) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  for (int i = tid; i < 9 * CC; i += DW_THREADS) {
    const int t9 = i / CC, ch = i - t9 * CC;
    s_w[t9][ch] = (c0 + ch < g.Cm) ? wgt[(c0 + ch) * 9 + t9] : 0.f;
  }
  if (tid == 0) {
    ptx::tma_prefetch_desc(&tm_in);
    ptx::mbar_init(&bar[0], 1);
    ptx::fence_barrier_init();
  }
  float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();

  const int TW2 = g.TW2, TH2 = g.TH2;
  const int npos = g.NI * TH2 * TW2;
  const uint32_t tx_bytes = (uint32_t)npos * CC * sizeof(T);
  const int nstrips = (g.TH + R - 1) / R;
  const DwItems im = dw_items<NTV>(g.TW);
  const int nrowitems = g.NI * nstrips;               // (image, strip) pairs of a tile
  const int nitems = nrowitems * g.TW * NTV;          // generic mapping
  const int x_fixed = (tid % im.lanes_per_row) / NTV;
  const int sub = tid / im.lanes_per_row;

  int t = worker;
  if (tid == 0 && t < g.ntiles) {
    int b0, h0, w0;
    dw_decode_tile(g, t, b0, h0, w0);
    ptx::mbar_arrive_expect_tx(&bar[0], tx_bytes);
    ptx::tma_load_4d(raw, &tm_in, &bar[0], c0, w0 - 1, h0 - 1, b0);
  }
  for (int it = 0; t < g.ntiles; t += g.nworkers, ++it) {
    int b0, h0, w0;
    dw_decode_tile(g, t, b0, h0, w0);
    ptx::mbar_wait(&bar[0], it & 1);

    // ---- BN1 + activation ONCE per element: raw (T) -> fp32 tile; positions outside the image -> 0 ----
    {
      float sc[AV], sh[AV];
#pragma unroll
      for (int k = 0; k < AV; ++k) { sc[k] = 0.f; sh[k] = 0.f; }
      if (cavalid) {
        ldv<AV>(scale1 + ca, sc);
        ldv<AV>(shift1 + ca, sh);
      }
#pragma unroll 2
      for (int pos = tid / NAV; pos < npos; pos += DW_THREADS / NAV) {
        const int r = fdiv(pos, g.d_tw2);
        const int col = pos - r * TW2;
        const int img = fdiv(r, g.d_th2);
        const int hr = r - img * TH2;
        const int gh = h0 - 1 + hr, gw = w0 - 1 + col;
        const bool inside = gh >= 0 && gh < g.H && gw >= 0 && gw < g.W && b0 + img < g.B && cavalid;
        float v[AV];
        ldv<AV>(raw + pos * CC + av * AV, v);
#pragma unroll
        for (int k = 0; k < AV; ++k) v[k] = inside ? ((g.dbg & 1) ? v[k] : act_apply_t<ACT, FastAct<T>::value>(fmaf(v[k], sc[k], sh[k]))) : 0.f;
        float* dst = tile + pos * CC + av * AV;
#pragma unroll
        for (int k = 0; k < AV; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    }
    ptx::fence_proxy_async();  // generic-proxy reads of `raw` precede the next TMA write into it
    __syncthreads();
    {
      const int tn = t + g.nworkers;
      if (tid == 0 && tn < g.ntiles) {
        int nb0, nh0, nw0;
        dw_decode_tile(g, tn, nb0, nh0, nw0);
        ptx::mbar_arrive_expect_tx(&bar[0], tx_bytes);
        ptx::tma_load_4d(raw, &tm_in, &bar[0], c0, nw0 - 1, nh0 - 1, nb0);
      }
    }

    // ---- 3x3 stencil: item = (img, strip of R rows, column x, 4-channel group) ----
    if (cvalid && !(g.dbg & 2)) {
      float w[9][4];
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9) {
        const float4 wv = *reinterpret_cast<const float4*>(&s_w[t9][tv * 4]);
        w[t9][0] = wv.x; w[t9][1] = wv.y; w[t9][2] = wv.z; w[t9][3] = wv.w;
      }
      const int step = im.regular ? im.rows_par : DW_THREADS;
      const int last = im.regular ? nrowitems : nitems;
      for (int s = im.regular ? sub : tid; s < last; s += step) {
        int x, rowitem;
        if (im.regular) {
          x = x_fixed;
          rowitem = s;
        } else {
          const int rest = s / NTV;
          rowitem = fdiv(rest, g.d_tw);
          x = rest - rowitem * g.TW;
        }
        const int img = fdiv(rowitem, g.d_ns4);
        const int r0 = (rowitem - img * nstrips) * R;
        const int b = b0 + img;
        if (b >= g.B || w0 + x >= g.W) continue;
        float acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[r][k] = 0.f;
        const float* base = tile + ((img * TH2 + r0) * TW2 + x) * CC + tv * 4;
#pragma unroll
        for (int ry = 0; ry < R + 2; ++ry) {  // input row (halo coordinates) r0 + ry
          if (r0 + ry >= TH2) break;
          float in[3][4];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4 q = *reinterpret_cast<const float4*>(base + (ry * TW2 + dx) * CC);
            in[dx][0] = q.x; in[dx][1] = q.y; in[dx][2] = q.z; in[dx][3] = q.w;
          }
#pragma unroll
          for (int ki = 0; ki < 3; ++ki) {
            const int r = ry - ki;  // output row of the strip that sees this input row through tap row ki
            if (r >= 0 && r < R) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[r][k] = fmaf(w[ki * 3 + dx][k], in[dx][k], acc[r][k]);
            }
          }
        }
        T* out = d_pre + (((long long)b * g.H + h0 + r0) * g.W + (w0 + x)) * g.Cm + c;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r0 + r < g.TH && h0 + r0 + r < g.H) {
            if (!(g.dbg & 4)) stv<4>(out + (long long)r * g.W * g.Cm, acc[r]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float sv = round_to<T>(acc[r][k]);  // statistics of the values as stored
              st_s[k] += sv;
              st_q[k] = fmaf(sv, sv, st_q[k]);
            }
          }
        }
      }
    }
    __syncthreads();  // fp32 tile is rewritten by the next step
  }
  if (cvalid && sum2) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&s_sum[tv * 4 + k], st_s[k]);
      atomicAdd(&s_sq[tv * 4 + k], st_q[k]);
    }
  }
  __syncthreads();
  if (tid < CC && c0 + tid < g.Cm) {
    if (sum2) atomicAdd(sum2 + c0 + tid, s_sum[tid]);
    if (sumsq2) atomicAdd(sumsq2 + c0 + tid, s_sq[tid]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: de_act[p] = sum_t w[t] * G[p - d_t];  du1 = de_act * act'(u1[p]);
//           dw[c,t] += sum_p e_act[p] * G[p - d_t];  dbeta1 += du1; dgamma1 += du1 * xhat1
// G = dd_pre (halo tile), e_pre centre tile -> u1 = scale1*e_pre + shift1.
// smem: raw G slot (T) | fp32 G tile | raw E slot (T) | barriers.  Per tile:
//   wait G -> convert G to fp32 -> sync -> TMA(next G) -> wait E -> stencil -> sync -> TMA(next E)
// ------------------------------------------------------------------------------------------------
template <typename T, int CC, int ACT, int TWc, int THc, int NIc>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_bwd_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_e,
                  const float* __restrict__ scale1, const float* __restrict__ shift1, const float* __restrict__ mean1,
                  const float* __restrict__ rstd1, const float* __restrict__ wgt, T* __restrict__ du1,
                  float* __restrict__ dwgt, float* __restrict__ dgamma1, float* __restrict__ dbeta1, const DwGeom g) {
  constexpr int R = 2;
  // TWc > 0: the tile geometry (columns, rows, images per tile) is a compile-time constant and every tile is full
  // (W % TW == 0, H % TH == 0, TH % R == 0): shared-memory offsets fold into immediates, the strip decode into shifts,
  // the row / column validity tests disappear.  ncu on the generic kernel: 60 issued instructions per element for 9
  // FFMA2, a third of them IMAD / IADD3 / SHF / ISETP / BRA from exactly that bookkeeping.
  constexpr bool SPEC = TWc > 0;
  constexpr int NTV = CC / 4;
  constexpr int AV = 16 / sizeof(T);
  constexpr int NAV = CC / AV;
  constexpr int BUF_ELEMS = DW_BUF_POS * CC;
  extern __shared__ __align__(128) uint8_t dsm[];
  // raw gradient halo slot | fp32 gradient tile | TWO pre-activation centre slots (the load of tile i+1 is issued at the
  // top of tile i, so its DRAM latency hides behind the whole tile instead of being waited for after the stencil)
  constexpr int E_ELEMS = 256 * CC;  // centre tile: TH * TW * NI <= 8 * 32 positions
  T* const graw = reinterpret_cast<T*>(dsm);
  float* const gt = reinterpret_cast<float*>(dsm + BUF_ELEMS * sizeof(T));
  T* const eraw0 = reinterpret_cast<T*>(dsm + BUF_ELEMS * (sizeof(T) + sizeof(float)));
  uint64_t* const bar = reinterpret_cast<uint64_t*>(dsm + BUF_ELEMS * (sizeof(T) + sizeof(float)) + 2 * E_ELEMS * sizeof(T));
  __shared__ float s_dw[CC * 9], s_db[CC], s_dg[CC];
  __shared__ __align__(16) float s_par[4][CC];  // scale1, shift1, mean1, rstd1 of the chunk
  __shared__ __align__(16) float s_w[9][CC];

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % g.nchunks;
  const int worker = blockIdx.x / g.nchunks;
  const int c0 = chunk * CC;
  const int tv = tid % NTV;
  const int c = c0 + tv * 4;
  const bool cvalid = c < g.Cm;
  const int av = tid % NAV;

  for (int i = tid; i < CC * 9; i += DW_THREADS) {
    s_dw[i] = 0.f;
    const int t9 = i / CC, ch = i - t9 * CC;
    s_w[t9][ch] = (c0 + ch < g.Cm) ? wgt[(c0 + ch) * 9 + t9] : 0.f;
  }
  if (tid < CC) {
    s_db[tid] = 0.f;
    s_dg[tid] = 0.f;
    const bool ok = c0 + tid < g.Cm;
    s_par[0][tid] = ok ? scale1[c0 + tid] : 0.f;
    s_par[1][tid] = ok ? shift1[c0 + tid] : 0.f;
    s_par[2][tid] = ok ? mean1[c0 + tid] : 0.f;
    s_par[3][tid] = ok ? rstd1[c0 + tid] : 0.f;
  }
  if (tid == 0) {
    ptx::tma_prefetch_desc(&tm_g);
    ptx::tma_prefetch_desc(&tm_e);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::mbar_init(&bar[2], 1);
    ptx::fence_barrier_init();
  }
  // channel pairs (k, k+1) ride in one 64-bit register: the 18 FMAs per element issue as 9 FFMA2
  // (measured -7 % on this kernel; the same change on the forward sweep was +3 %: its pack moves outweigh the gain)
  f32x2 dwa[9][2];
#pragma unroll
  for (int t9 = 0; t9 < 9; ++t9) dwa[t9][0] = dwa[t9][1] = 0ull;
  float a_db[4] = {0.f, 0.f, 0.f, 0.f}, a_dg[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();

  const int TW = SPEC ? TWc : g.TW, TH = SPEC ? THc : g.TH, NI = SPEC ? NIc : g.NI;
  const int TW2 = TW + 2, TH2 = TH + 2;
  const int npos = NI * TH2 * TW2;
  const uint32_t g_bytes = (uint32_t)npos * CC * sizeof(T);
  const uint32_t e_bytes = (uint32_t)(NI * TH * TW) * CC * sizeof(T);
  const int nstrips = (TH + R - 1) / R;
  const DwItems im = dw_items<NTV>(TW);
  const int nrowitems = NI * nstrips;
  const int nitems = nrowitems * TW * NTV;
  const int x_fixed = (tid % im.lanes_per_row) / NTV;
  const int sub = tid / im.lanes_per_row;

  int t = worker;
  if (tid == 0 && t < g.ntiles) {
    int b0, h0, w0;
    dw_decode_tile(g, t, b0, h0, w0);
    ptx::mbar_arrive_expect_tx(&bar[0], g_bytes);
    ptx::tma_load_4d(graw, &tm_g, &bar[0], c0, w0 - 1, h0 - 1, b0);
    ptx::mbar_arrive_expect_tx(&bar[1], e_bytes);
    ptx::tma_load_4d(eraw0, &tm_e, &bar[1], c0, w0, h0, b0);
  }
  for (int it = 0; t < g.ntiles; t += g.nworkers, ++it) {
    int b0, h0, w0;
    dw_decode_tile(g, t, b0, h0, w0);
    const int tn = t + g.nworkers;
    int nb0 = 0, nh0 = 0, nw0 = 0;
    if (tn < g.ntiles) dw_decode_tile(g, tn, nb0, nh0, nw0);
    const int eb = it & 1;
    const T* const eraw = eraw0 + eb * E_ELEMS;
    // the other centre slot was last read by tile it-1's stencil, which every thread left before the closing barrier
    if (tid == 0 && tn < g.ntiles) {
      ptx::mbar_arrive_expect_tx(&bar[1 + (eb ^ 1)], e_bytes);
      ptx::tma_load_4d(eraw0 + (eb ^ 1) * E_ELEMS, &tm_e, &bar[1 + (eb ^ 1)], c0, nw0, nh0, nb0);
    }

    // ---- gradient halo tile: raw (T) -> fp32 (TMA zero fill already handled the borders) ----
    ptx::mbar_wait(&bar[0], it & 1);
    // 4 elements per thread: 8-byte loads / 16-byte stores with consecutive lanes on consecutive addresses,
    // both conflict-free (8 elements per thread made every float4 store 2-way bank conflicted)
#pragma unroll 4
    for (int i = tid; i < npos * NTV; i += DW_THREADS) {
      float v[4];
      ldv<4>(graw + i * 4, v);
      *reinterpret_cast<float4*>(gt + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    (void)av;
    (void)NAV;
    ptx::fence_proxy_async();
    __syncthreads();
    if (tid == 0 && tn < g.ntiles) {
      ptx::mbar_arrive_expect_tx(&bar[0], g_bytes);
      ptx::tma_load_4d(graw, &tm_g, &bar[0], c0, nw0 - 1, nh0 - 1, nb0);
    }
    ptx::mbar_wait(&bar[1 + eb], (it >> 1) & 1);

    if (cvalid) {
      f32x2 w[9][2];
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9) {
        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(&s_w[t9][tv * 4]);
        w[t9][0] = wv.x; w[t9][1] = wv.y;
      }
      const int step = im.regular ? im.rows_par : DW_THREADS;
      const int last = im.regular ? nrowitems : nitems;
#pragma unroll 1
      for (int s = im.regular ? sub : tid; s < last; s += step) {
        int x, rowitem;
        if (im.regular) {
          x = x_fixed;
          rowitem = s;
        } else {
          const int rest = s / NTV;
          rowitem = fdiv(rest, g.d_tw);
          x = rest - rowitem * TW;
        }
        const int img = SPEC ? rowitem / nstrips : fdiv(rowitem, g.d_ns2);
        const int r0 = (rowitem - img * nstrips) * R;
        const int b = b0 + img;
        if (b >= g.B || (!SPEC && w0 + x >= g.W)) continue;
        f32x2 de[R][2], ea[R][2];
        float da[R][4];
        bool rvalid[R];
        const T* ep = eraw + ((img * TH + r0) * TW + x) * CC + tv * 4;
        {
          const float4 sc = *reinterpret_cast<const float4*>(&s_par[0][tv * 4]);
          const float4 sh = *reinterpret_cast<const float4*>(&s_par[1][tv * 4]);
          const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
          for (int r = 0; r < R; ++r) {
            rvalid[r] = SPEC || ((r0 + r < TH) && (h0 + r0 + r < g.H));
            de[r][0] = de[r][1] = ea[r][0] = ea[r][1] = 0ull;
#pragma unroll
            for (int k = 0; k < 4; ++k) da[r][k] = 0.f;
            if (rvalid[r]) {
              float ev[4], av4[4];
              ldv<4>(ep + r * TW * CC, ev);
#pragma unroll
              for (int k = 0; k < 4; ++k) act_both_t<ACT, FastAct<T>::value>(fmaf(ev[k], scv[k], shv[k]), &av4[k], &da[r][k]);
              ea[r][0] = pk2(av4[0], av4[1]);
              ea[r][1] = pk2(av4[2], av4[3]);
            }
          }
        }
        const float* base = gt + ((img * TH2 + r0) * TW2 + x) * CC + tv * 4;
#pragma unroll
        for (int ry = 0; ry < R + 2; ++ry) {  // gradient row (halo coords) r0 + ry  <->  image row h0 + r0 + ry - 1
          if (!SPEC && r0 + ry >= TH2) break;
          ulonglong2 gr[3];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) gr[dx] = *reinterpret_cast<const ulonglong2*>(base + (ry * TW2 + dx) * CC);
          // output row r sits at halo row r0 + r + 1; gradient row offset (ry - 1) - r = -(ki - 1)
#pragma unroll
          for (int ki = 0; ki < 3; ++ki) {
            const int r = ry - 1 + (ki - 1);
            if (r >= 0 && r < R) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int tt = ki * 3 + (2 - dx);  // gradient column offset dx - 1 = -(kj - 1)
                de[r][0] = fma2(w[tt][0], gr[dx].x, de[r][0]);
                de[r][1] = fma2(w[tt][1], gr[dx].y, de[r][1]);
                dwa[tt][0] = fma2(ea[r][0], gr[dx].x, dwa[tt][0]);
                dwa[tt][1] = fma2(ea[r][1], gr[dx].y, dwa[tt][1]);
              }
            }
          }
        }
        T* out = du1 + (((long long)b * g.H + h0 + r0) * g.W + (w0 + x)) * g.Cm + c;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (rvalid[r]) {
            float o[4], ev[4], dv[4];
            ldv<4>(ep + r * TW * CC, ev);  // re-read (smem) instead of carrying it through the stencil
            unpk2(de[r][0], dv[0], dv[1]);
            unpk2(de[r][1], dv[2], dv[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              o[k] = dv[k] * da[r][k];
              a_db[k] += o[k];
              a_dg[k] = fmaf(o[k], ev[k], a_dg[k]);  // sum o*e; xhat = (e - mean)*rstd is folded in at the end
            }
            stv<4>(out + (long long)r * g.W * g.Cm, o);
          }
        }
      }
    }
    ptx::fence_proxy_async();
    __syncthreads();
  }
  if (cvalid) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&s_db[tv * 4 + k], a_db[k]);
      // sum o*xhat = rstd * (sum o*e - mean * sum o), per thread (a few hundred terms) before the CTA fold
      atomicAdd(&s_dg[tv * 4 + k], s_par[3][tv * 4 + k] * (a_dg[k] - s_par[2][tv * 4 + k] * a_db[k]));
    }
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {
      float d4[4];
      unpk2(dwa[t9][0], d4[0], d4[1]);
      unpk2(dwa[t9][1], d4[2], d4[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&s_dw[(tv * 4 + k) * 9 + t9], d4[k]);
    }
  }
  __syncthreads();
  for (int i = tid; i < CC * 9; i += DW_THREADS)
    if (c0 + i / 9 < g.Cm) atomicAdd(dwgt + (long long)c0 * 9 + i, s_dw[i]);
  if (tid < CC && c0 + tid < g.Cm) {
    atomicAdd(dbeta1 + c0 + tid, s_db[tid]);
    atomicAdd(dgamma1 + c0 + tid, s_dg[tid]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward, register-window variant.  lane = 4 consecutive channels, warp = 128 channels of one
// (image, band of RR output rows); the warp sweeps x = 0..W-1 holding the 3 x (RR+2) x 4 window of
// ACTIVATED inputs in registers, so there is no shared-memory staging, no barrier and no tile
// bookkeeping: every step is (RR+2) coalesced 256-byte row loads, (RR+2)*4 activations, RR*36 FMAs and
// RR coalesced stores.  Halo rows are re-read by the neighbouring band (same CTA -> L1/L2 hits).
// ------------------------------------------------------------------------------------------------
template <typename T, int ACT, int RR>
__global__ void __launch_bounds__(256, 2)
dwconv_fwd_sweep_kernel(const T* __restrict__ e_pre, const float* __restrict__ scale1, const float* __restrict__ shift1,
                        const float* __restrict__ wgt, T* __restrict__ d_pre, float* __restrict__ sum2,
                        float* __restrict__ sumsq2, int B, int H, int W, int Cm, int bands_per_img, int nbands) {
  constexpr int NR = RR + 2;
  __shared__ float s_red[2][8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + lane) * 4;
  const bool cvalid = c < Cm;
  float w[9][4], sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sc[k] = cvalid ? scale1[c + k] : 0.f;
    sh[k] = cvalid ? shift1[c + k] : 0.f;
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) w[t9][k] = cvalid ? wgt[(c + k) * 9 + t9] : 0.f;
  }
  float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
  const long long rowpitch = (long long)W * Cm;

  for (int item = blockIdx.x * 8 + warp; item < nbands && cvalid; item += gridDim.x * 8) {
    const int b = item / bands_per_img;
    const int r0 = (item - b * bands_per_img) * RR;
    const T* in0 = e_pre + ((long long)b * H + (r0 - 1)) * rowpitch + c;   // halo row 0, column 0
    T* out0 = d_pre + ((long long)b * H + r0) * rowpitch + c;
    bool rv[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) rv[i] = (r0 - 1 + i >= 0) && (r0 - 1 + i < H);
    float win[3][NR][4];
    // load + activate one column of the window (zeros outside the image)
    auto load_col = [&](float (&col)[NR][4], int x) {
      const bool xv = x < W;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const bool ok = xv && rv[i];
        if (ok) ldv<4>(in0 + (long long)i * rowpitch + (long long)x * Cm, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) col[i][k] = ok ? act_apply_t<ACT, FastAct<T>::value>(fmaf(v[k], sc[k], sh[k])) : 0.f;
      }
    };
    auto step = [&](const float (&L)[NR][4], const float (&M)[NR][4], const float (&Rc)[NR][4], int x) {
      float acc[RR][4];
#pragma unroll
      for (int r = 0; r < RR; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a = 0.f;
#pragma unroll
          for (int ki = 0; ki < 3; ++ki) {
            a = fmaf(w[ki * 3 + 0][k], L[r + ki][k], a);
            a = fmaf(w[ki * 3 + 1][k], M[r + ki][k], a);
            a = fmaf(w[ki * 3 + 2][k], Rc[r + ki][k], a);
          }
          acc[r][k] = a;
        }
#pragma unroll
      for (int r = 0; r < RR; ++r) {
        if (r0 + r < H) {
          stv<4>(out0 + (long long)r * rowpitch + (long long)x * Cm, acc[r]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // statistics of the values AS STORED: BatchNorm's backward sums (sum xhat = 0, sum xhat^2 = n) only
            // cancel exactly if mean / variance describe the tensor the later passes actually read
            const float sv = round_to<T>(acc[r][k]);
            st_s[k] += sv;
            st_q[k] = fmaf(sv, sv, st_q[k]);
          }
        }
      }
    };
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) win[0][i][k] = 0.f;
    load_col(win[1], 0);
    for (int x = 0; x < W; x += 3) {
      load_col(win[2], x + 1);
      step(win[0], win[1], win[2], x);
      if (x + 1 < W) {
        load_col(win[0], x + 2);
        step(win[1], win[2], win[0], x + 1);
      }
      if (x + 2 < W) {
        load_col(win[1], x + 3);
        step(win[2], win[0], win[1], x + 2);
      }
    }
  }
  // per-channel statistics: fold the 8 warps of the CTA, then one atomic per channel per CTA
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s_red[0][warp][lane * 4 + k] = st_s[k];
    s_red[1][warp][lane * 4 + k] = st_q[k];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int ch = blockIdx.y * 128 + threadIdx.x;
    if (ch < Cm) {
      float a = 0.f, q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { a += s_red[0][j][threadIdx.x]; q += s_red[1][j][threadIdx.x]; }
      if (sum2) atomicAdd(sum2 + ch, a);
      if (sumsq2) atomicAdd(sumsq2 + ch, q);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward, register-window sweep fed by a per-warp cp.async ring.  The sweep above issues the loads of window column
// x + 1 and needs them at once: one column (4 rows x 256 B per warp) in flight per warp, 16 KB per SM -- half of what
// the HBM latency-bandwidth product asks for, and `long_scoreboard` is its top stall.  Here every warp streams its
// columns through a private shared-memory ring KD columns deep (cp.async, 8/16 bytes per lane, no barrier: a lane
// only ever reads back the bytes it copied itself), so KD columns per warp are in flight, across band boundaries too
// (the prefetch cursor runs ahead into the warp's next band).  Registers, math and stores are those of the sweep.
// ------------------------------------------------------------------------------------------------
template <typename T, int ACT, int RR, int KD>
__global__ void __launch_bounds__(256, 2)
dwconv_fwd_sweep_pf_kernel(const T* __restrict__ e_pre, const float* __restrict__ scale1, const float* __restrict__ shift1,
                           const float* __restrict__ wgt, T* __restrict__ d_pre, float* __restrict__ sum2,
                           float* __restrict__ sumsq2, int B, int H, int W, int Cm, int bands_per_img, int nbands) {
  constexpr int NR = RR + 2;
  constexpr int VB = 4 * (int)sizeof(T);  // bytes per lane and (row, column): 4 channels
  extern __shared__ __align__(16) uint8_t dw_ring[];  // [8 warps][KD][NR][32 lanes][VB]
  __shared__ float s_red[2][8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + lane) * 4;
  const bool cvalid = c < Cm;
  uint8_t* const ring = dw_ring + (size_t)warp * KD * NR * 32 * VB + lane * VB;
  float w[9][4], sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sc[k] = cvalid ? scale1[c + k] : 0.f;
    sh[k] = cvalid ? shift1[c + k] : 0.f;
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) w[t9][k] = cvalid ? wgt[(c + k) * 9 + t9] : 0.f;
  }
  float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
  const long long rowpitch = (long long)W * Cm;
  const int item0 = blockIdx.x * 8 + warp, istride = gridDim.x * 8;

  // ---- prefetch cursor: (band, column) of the next column to request ----
  int pf_item = cvalid ? item0 : nbands, pf_n = 0;
  const T* pf_in0 = e_pre;
  unsigned pf_rv = 0;
  auto pf_setup = [&]() {
    if (pf_item < nbands) {
      const int b = pf_item / bands_per_img;
      const int r0 = (pf_item - b * bands_per_img) * RR;
      pf_in0 = e_pre + ((long long)b * H + (r0 - 1)) * rowpitch + c;
      pf_rv = 0;
#pragma unroll
      for (int i = 0; i < NR; ++i) pf_rv |= ((r0 - 1 + i >= 0) && (r0 - 1 + i < H)) ? (1u << i) : 0u;
    }
  };
  auto prefetch = [&](int slot) {
    if (pf_item < nbands) {
      if (pf_n < W) {
#pragma unroll
        for (int i = 0; i < NR; ++i)
          if ((pf_rv >> i) & 1u)
            ptx::cp_async<VB>(ring + (size_t)(slot * NR + i) * 32 * VB, pf_in0 + (long long)i * rowpitch + (long long)pf_n * Cm);
      }
      if (++pf_n > W) {  // W + 1 columns per band: 0 .. W-1 and the all-zero column beyond the right edge
        pf_n = 0;
        pf_item += istride;
        pf_setup();
      }
    }
    ptx::cp_async_commit();  // one group per column, empty or not: the consumer counts groups
  };
  pf_setup();
#pragma unroll
  for (int k = 0; k < KD; ++k) prefetch(k);
  int cs = 0;  // ring slot of the next column to consume

  for (int item = item0; item < nbands && cvalid; item += istride) {
    const int b = item / bands_per_img;
    const int r0 = (item - b * bands_per_img) * RR;
    T* out0 = d_pre + ((long long)b * H + r0) * rowpitch + c;
    bool rv[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) rv[i] = (r0 - 1 + i >= 0) && (r0 - 1 + i < H);
    float win[3][NR][4];
    // take one column of the window out of the ring + activate it (zeros outside the image), refill the slot
    auto load_col = [&](float (&col)[NR][4], int x) {
      ptx::cp_async_wait<KD - 1>();
      const bool xv = x < W;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const bool ok = xv && rv[i];
        if (ok) ldv<4>(reinterpret_cast<const T*>(ring + (size_t)(cs * NR + i) * 32 * VB), v);
#pragma unroll
        for (int k = 0; k < 4; ++k) col[i][k] = ok ? act_apply_t<ACT, FastAct<T>::value>(fmaf(v[k], sc[k], sh[k])) : 0.f;
      }
      prefetch(cs);
      cs = cs + 1 == KD ? 0 : cs + 1;
    };
    auto step = [&](const float (&L)[NR][4], const float (&M)[NR][4], const float (&Rc)[NR][4], int x) {
      float acc[RR][4];
#pragma unroll
      for (int r = 0; r < RR; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a = 0.f;
#pragma unroll
          for (int ki = 0; ki < 3; ++ki) {
            a = fmaf(w[ki * 3 + 0][k], L[r + ki][k], a);
            a = fmaf(w[ki * 3 + 1][k], M[r + ki][k], a);
            a = fmaf(w[ki * 3 + 2][k], Rc[r + ki][k], a);
          }
          acc[r][k] = a;
        }
#pragma unroll
      for (int r = 0; r < RR; ++r) {
        if (r0 + r < H) {
          stv<4>(out0 + (long long)r * rowpitch + (long long)x * Cm, acc[r]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float sv = round_to<T>(acc[r][k]);  // statistics of the values AS STORED (see the sweep above)
            st_s[k] += sv;
            st_q[k] = fmaf(sv, sv, st_q[k]);
          }
        }
      }
    };
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) win[0][i][k] = 0.f;
    load_col(win[1], 0);
    for (int x = 0; x < W; x += 3) {
      load_col(win[2], x + 1);
      step(win[0], win[1], win[2], x);
      if (x + 1 < W) {
        load_col(win[0], x + 2);
        step(win[1], win[2], win[0], x + 1);
      }
      if (x + 2 < W) {
        load_col(win[1], x + 3);
        step(win[2], win[0], win[1], x + 2);
      }
    }
  }
  ptx::cp_async_wait<0>();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s_red[0][warp][lane * 4 + k] = st_s[k];
    s_red[1][warp][lane * 4 + k] = st_q[k];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int ch = blockIdx.y * 128 + threadIdx.x;
    if (ch < Cm) {
      float a = 0.f, q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { a += s_red[0][j][threadIdx.x]; q += s_red[1][j][threadIdx.x]; }
      if (sum2) atomicAdd(sum2 + ch, a);
      if (sumsq2) atomicAdd(sumsq2 + ch, q);
    }
  }
}

// opt in to the dynamic shared memory and report how many CTAs of this kernel are resident per SM:
// the persistent grid is sized to exactly one wave (a second, partly filled wave would idle SMs).
template <typename K>
int dw_smem_optin(K kernel, int bytes, int* ctas_per_sm) {
  // keyed by the kernel ADDRESS: instantiations that share a signature share this function (and its statics)
  struct Entry { const void* fn; int bytes, occ; };
  static Entry cache[32] = {};
  static int used = 0;
  const void* fn = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < used; ++i)
    if (cache[i].fn == fn && cache[i].bytes == bytes) { *ctas_per_sm = cache[i].occ; return OGV_OK; }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    ogv_set_error("dwconv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return OGV_ERR_CUDA;
  }
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, DW_THREADS, bytes);
  if (e != cudaSuccess || occ < 1) occ = 1;
  if (used < 32) cache[used++] = Entry{fn, bytes, occ};
  *ctas_per_sm = occ;
  return OGV_OK;
}

template <typename T>
constexpr int dw_cc() { return 64 / (int)sizeof(T); }  // 64-byte channel chunk per position

// ------------------------------------------------------------------------------------------------
// forward, TMA tile + register-window walkers (bf16, W in {4, 8, 16, 32}, Cm % 64 == 0: every stage of the model
// configs).  A persistent CTA owns a 64-channel chunk and walks (image group, row band) tiles; the halo tile of step
// i+1 is in flight (cp.async.bulk.tensor.4d, zero-filled borders) while step i is computed, so no thread ever waits
// for DRAM -- the sweep kernel above is latency-bound at 16 warps per SM.  Every thread is a walker: two adjacent
// columns x four channels, moving down the band with the 3 x 4 window of ACTIVATED inputs in registers: one new row
// (4 LDS.64, 16 BN-affine + activation) per 8 outputs, tap weights in registers, 36 FFMA2 per step.
// ------------------------------------------------------------------------------------------------
constexpr int DWW_CC = 64;
template <int W_> struct DwwShape;
template <> struct DwwShape<32> { static constexpr int NI = 1, NSB = 1, TRS = 8; };
template <> struct DwwShape<16> { static constexpr int NI = 1, NSB = 2, TRS = 8; };
template <> struct DwwShape<8> { static constexpr int NI = 4, NSB = 1, TRS = 8; };
template <> struct DwwShape<4> { static constexpr int NI = 8, NSB = 1, TRS = 4; };

struct DwwGeom {
  int B, H, W, Cm, bands, tiles_w, ntiles, nchunks, nworkers;  // W = IMAGE width = tiles_w x the tile width
};

template <int ACT, int W_>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_fwd_walk_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ scale1,
                       const float* __restrict__ shift1, const float* __restrict__ wgt, bf16* __restrict__ d_pre,
                       float* __restrict__ sum2, float* __restrict__ sumsq2, const DwwGeom g) {
  using S = DwwShape<W_>;
  constexpr int NI = S::NI, NSB = S::NSB, TRS = S::TRS, TRT = TRS * NSB;
  constexpr int W2 = W_ + 2, TR2 = TRT + 2;
  constexpr int TILE_ELEMS = NI * TR2 * W2 * DWW_CC;
  constexpr int ROW = W2 * DWW_CC;
  extern __shared__ __align__(128) uint8_t dsm[];
  bf16* const slot0 = reinterpret_cast<bf16*>(dsm);
  __shared__ uint64_t bar[2];
  __shared__ float s_sum[DWW_CC], s_sq[DWW_CC];

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % g.nchunks;
  const int worker = blockIdx.x / g.nchunks;
  const int c0 = chunk * DWW_CC;
  const int cg = tid % 16;
  const int xp = (tid / 16) % (W_ / 2);
  const int grp = tid / (8 * W_);
  const int img = grp / NSB, rs = (grp % NSB) * TRS;
  const int x0 = xp * 2;
  const int c = c0 + cg * 4;

  if (tid < DWW_CC) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
  if (tid == 0) {
    ptx::tma_prefetch_desc(&tm_in);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  // tap weights and BN1 affine of the thread's four channels, as channel pairs
  f32x2 w2[9][2], sc2[2], sh2[2];
  {
    float wv[4][9];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) wv[k][t] = wgt[(long long)(c + k) * 9 + t];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      w2[t][0] = pk2(wv[0][t], wv[1][t]);
      w2[t][1] = pk2(wv[2][t], wv[3][t]);
    }
    const float4 sc = *reinterpret_cast<const float4*>(scale1 + c);
    const float4 sh = *reinterpret_cast<const float4*>(shift1 + c);
    sc2[0] = pk2(sc.x, sc.y); sc2[1] = pk2(sc.z, sc.w);
    sh2[0] = pk2(sh.x, sh.y); sh2[1] = pk2(sh.z, sh.w);
  }
  float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};
  const int Wi = g.W;
  __syncthreads();

  // tile t = ((image group * bands) + band) * tiles_w + column tile (images wider than the tile are cut in columns)
  auto tile_coords = [&](int t, int& b0, int& r0, int& w0) {
    const int q = t / g.tiles_w;
    w0 = (t - q * g.tiles_w) * W_;
    const int bgrp = q / g.bands;
    b0 = bgrp * NI;
    r0 = (q - bgrp * g.bands) * TRT;
  };
  int t = worker;
  if (tid == 0 && t < g.ntiles) {
    int b0, r0, w0;
    tile_coords(t, b0, r0, w0);
    ptx::mbar_arrive_expect_tx(&bar[0], (uint32_t)(TILE_ELEMS * sizeof(bf16)));
    ptx::tma_load_4d(slot0, &tm_in, &bar[0], c0, w0 - 1, r0 - 1, b0);
  }
  for (int it = 0; t < g.ntiles; t += g.nworkers, ++it) {
    int b0, r0, w0;
    tile_coords(t, b0, r0, w0);
    const int sl = it & 1;
    const int tn = t + g.nworkers;
    // the other slot was last read in iteration it-1, which every thread left through the closing barrier
    if (tid == 0 && tn < g.ntiles) {
      int nb0, nr0, nw0;
      tile_coords(tn, nb0, nr0, nw0);
      ptx::mbar_arrive_expect_tx(&bar[sl ^ 1], (uint32_t)(TILE_ELEMS * sizeof(bf16)));
      ptx::tma_load_4d(slot0 + (sl ^ 1) * TILE_ELEMS, &tm_in, &bar[sl ^ 1], c0, nw0 - 1, nr0 - 1, nb0);
    }
    ptx::mbar_wait(&bar[sl], (it >> 1) & 1);
    const int b = b0 + img;
    if (b < g.B) {
      // window columns are image columns w0 + x0 - 1 .. w0 + x0 + 2: only the outer two can fall off the IMAGE
      const bool col_ok[4] = {w0 + x0 >= 1, true, true, w0 + x0 + 2 < Wi};
      const bf16* tcol = slot0 + sl * TILE_ELEMS + ((img * TR2 + rs) * W2 + x0) * DWW_CC + cg * 4;
      const int h0 = r0 + rs;  // image row of the walker's first output; tile row j <-> image row h0 - 1 + j
      bf16* out = d_pre + (((long long)b * g.H + h0) * Wi + w0 + x0) * g.Cm + c;
      const long long row_stride = (long long)Wi * g.Cm;
      f32x2 win[3][4][2];
      auto load_row = [&](int j, f32x2 (&row)[4][2]) {
        const int h = h0 - 1 + j;
        const bool rok = h >= 0 && h < g.H;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const uint2 u = *reinterpret_cast<const uint2*>(tcol + j * ROW + cc * DWW_CC);
          const f32x2 x01 = pk2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
          const f32x2 x23 = pk2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
          float a0, a1, a2, a3;
          unpk2(fma2(x01, sc2[0], sh2[0]), a0, a1);
          unpk2(fma2(x23, sc2[1], sh2[1]), a2, a3);
          const bool ok = rok && col_ok[cc];  // outside the image the conv input is ZERO, not act(shift)
          a0 = ok ? act_apply_t<ACT, true>(a0) : 0.f;
          a1 = ok ? act_apply_t<ACT, true>(a1) : 0.f;
          a2 = ok ? act_apply_t<ACT, true>(a2) : 0.f;
          a3 = ok ? act_apply_t<ACT, true>(a3) : 0.f;
          row[cc][0] = pk2(a0, a1);
          row[cc][1] = pk2(a2, a3);
        }
      };
      load_row(0, win[0]);
      load_row(1, win[1]);
#pragma unroll
      for (int j = 0; j < TRS; ++j) {
        load_row(j + 2, win[(j + 2) % 3]);
        f32x2 acc[2][2] = {{0ull, 0ull}, {0ull, 0ull}};
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) {
          const int r = (j + t9 / 3) % 3;
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            acc[p][0] = fma2(w2[t9][0], win[r][t9 % 3 + p][0], acc[p][0]);
            acc[p][1] = fma2(w2[t9][1], win[r][t9 % 3 + p][1], acc[p][1]);
          }
        }
        if (h0 + j < g.H) {
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            float f[4];
            unpk2(acc[p][0], f[0], f[1]);
            unpk2(acc[p][1], f[2], f[3]);
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(f[0], f[1]), h23 = __floats2bfloat162_rn(f[2], f[3]);
            *reinterpret_cast<uint2*>(out + j * row_stride + p * g.Cm) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
            // statistics of the values as stored
            const float2 s01 = __bfloat1622float2(h01), s23 = __bfloat1622float2(h23);
            st_s[0] += s01.x; st_q[0] = fmaf(s01.x, s01.x, st_q[0]);
            st_s[1] += s01.y; st_q[1] = fmaf(s01.y, s01.y, st_q[1]);
            st_s[2] += s23.x; st_q[2] = fmaf(s23.x, s23.x, st_q[2]);
            st_s[3] += s23.y; st_q[3] = fmaf(s23.y, s23.y, st_q[3]);
          }
        }
      }
    }
    __syncthreads();  // every walker is done with slot `sl` before the next iteration's TMA overwrites it
  }
  if (sum2) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // fold the 16 / W_ * ... walkers that share a channel inside the warp first (lanes cg, cg + 16)
      float a = st_s[k] + __shfl_xor_sync(0xffffffffu, st_s[k], 16);
      float q = st_q[k] + __shfl_xor_sync(0xffffffffu, st_q[k], 16);
      if ((tid & 16) == 0) {
        atomicAdd(&s_sum[cg * 4 + k], a);
        atomicAdd(&s_sq[cg * 4 + k], q);
      }
    }
    __syncthreads();
    if (tid < DWW_CC) {
      atomicAdd(sum2 + c0 + tid, s_sum[tid]);
      if (sumsq2) atomicAdd(sumsq2 + c0 + tid, s_sq[tid]);
    }
  }
}

template <int ACT, int W_>
int dww_launch(const void* e_pre, const float* scale1, const float* shift1, const float* w, void* d_pre, float* sum2,
               float* sumsq2, int B, int H, int W, int Cm, cudaStream_t st) {
  using S = DwwShape<W_>;
  constexpr int TRT = S::TRS * S::NSB;
  constexpr int TILE_BYTES = S::NI * (TRT + 2) * (W_ + 2) * DWW_CC * 2;
  const int smem = 2 * TILE_BYTES;
  auto kern = dwconv_fwd_walk_kernel<ACT, W_>;
  int occ = 1;
  if (int rc = dw_smem_optin(kern, smem, &occ)) return rc;
  DwwGeom g;
  g.B = B; g.H = H; g.W = W; g.Cm = Cm;
  g.bands = (H + TRT - 1) / TRT;
  g.tiles_w = W / W_;
  const long long nt = (long long)((B + S::NI - 1) / S::NI) * g.bands * g.tiles_w;
  if (nt > 0x7fffffffLL) { ogv_set_error("dwconv_fwd: too many tiles"); return OGV_ERR_UNSUPPORTED; }
  g.ntiles = (int)nt;
  g.nchunks = Cm / DWW_CC;
  long long want = ((long long)ogv_num_sms() * occ) / g.nchunks;
  if (want > nt) want = nt;
  if (want < 1) want = 1;
  g.nworkers = (int)want;
  CUtensorMap tm;
  {
    unsigned long long dims[4] = {(unsigned long long)Cm, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
    unsigned long long str[3] = {(unsigned long long)Cm * 2, (unsigned long long)W * Cm * 2, (unsigned long long)H * W * Cm * 2};
    unsigned box[4] = {(unsigned)DWW_CC, (unsigned)(W_ + 2), (unsigned)(TRT + 2), (unsigned)S::NI};
    if (int rc = ogv_make_tmap(&tm, e_pre, OGV_BF16, 4, dims, str, box, 0)) return rc;
  }
  kern<<<g.nchunks * g.nworkers, DW_THREADS, smem, st>>>(tm, scale1, shift1, w, reinterpret_cast<bf16*>(d_pre), sum2,
                                                         sumsq2, g);
  return ogv_check_launch("dwconv_fwd");
}

// ------------------------------------------------------------------------------------------------
// backward, TMA tiles + register-window walkers (bf16, W in {4, 8, 16, 32}, Cm % 32 == 0).  Same structure as the
// forward walker; a thread owns two adjacent columns x TWO channels (one fp32 pair) so that the 9 tap weights, the 9
// filter-gradient accumulators and the 3 x 4 gradient window fit in registers together (four channels need 136).
// Per step: window element (i, k) = G[q + (i-1, k-1)] carries weight w[8 - (3i+k)] into de[q] and a[q] into
// dw[8 - (3i+k)]; du1 = de * act'(u1), dbeta1 += du1, dgamma1 += du1 * xhat (x-form, folded with mean / rstd at the end).
// The gradient tile needs no masks at all (TMA zero fill); rows / images outside the tensor are skipped per warp.
// ------------------------------------------------------------------------------------------------
constexpr int DWB_CC = 32;
#ifndef DWB_MINB
#define DWB_MINB 2
#endif

template <int ACT, int W_>
__global__ void __launch_bounds__(DW_THREADS, DWB_MINB)
dwconv_bwd_walk_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_e,
                       const float* __restrict__ scale1, const float* __restrict__ shift1,
                       const float* __restrict__ mean1, const float* __restrict__ rstd1, const float* __restrict__ wgt,
                       bf16* __restrict__ du1, float* __restrict__ dwgt, float* __restrict__ dgamma1,
                       float* __restrict__ dbeta1, const DwwGeom g) {
  using S = DwwShape<W_>;
  constexpr int NI = S::NI, NSB = S::NSB, TRS = S::TRS, TRT = TRS * NSB;
  constexpr int W2 = W_ + 2, TR2 = TRT + 2;
  constexpr int G_ELEMS = NI * TR2 * W2 * DWB_CC, E_ELEMS = NI * TRT * W_ * DWB_CC;
  constexpr int SLOT = G_ELEMS + E_ELEMS;
  constexpr int GROW = W2 * DWB_CC, EROW = W_ * DWB_CC;
  extern __shared__ __align__(128) uint8_t dsm[];
  bf16* const slot0 = reinterpret_cast<bf16*>(dsm);
  __shared__ uint64_t bar[2];
  __shared__ float s_dw[DWB_CC * 9], s_db[DWB_CC], s_dg[DWB_CC];

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % g.nchunks;
  const int worker = blockIdx.x / g.nchunks;
  const int c0 = chunk * DWB_CC;
  const int cg = tid % 16;
  const int xp = (tid / 16) % (W_ / 2);
  const int grp = tid / (8 * W_);
  const int img = grp / NSB, rs = (grp % NSB) * TRS;
  const int x0 = xp * 2;
  const int c = c0 + cg * 2;

  for (int i = tid; i < DWB_CC * 9; i += DW_THREADS) s_dw[i] = 0.f;
  if (tid < DWB_CC) { s_db[tid] = 0.f; s_dg[tid] = 0.f; }
  if (tid == 0) {
    ptx::tma_prefetch_desc(&tm_g);
    ptx::tma_prefetch_desc(&tm_e);
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  f32x2 wf[9];  // wf[s] = w[8 - s]: the weight window element s = 3i + k carries
#pragma unroll
  for (int s9 = 0; s9 < 9; ++s9) wf[s9] = pk2(wgt[(long long)c * 9 + 8 - s9], wgt[(long long)(c + 1) * 9 + 8 - s9]);
  const f32x2 sc2 = pk2(scale1[c], scale1[c + 1]), sh2 = pk2(shift1[c], shift1[c + 1]);
  f32x2 dwa[9];  // dwa[s] accumulates dw[8 - s]
#pragma unroll
  for (int s9 = 0; s9 < 9; ++s9) dwa[s9] = 0ull;
  f32x2 a_db = 0ull, a_dg = 0ull;
  __syncthreads();

  const int Wi = g.W;
  auto tile_coords = [&](int t, int& b0, int& r0, int& w0) {
    const int q = t / g.tiles_w;
    w0 = (t - q * g.tiles_w) * W_;
    const int bgrp = q / g.bands;
    b0 = bgrp * NI;
    r0 = (q - bgrp * g.bands) * TRT;
  };
  auto issue = [&](int t, int sl) {
    int b0, r0, w0;
    tile_coords(t, b0, r0, w0);
    bf16* dst = slot0 + sl * SLOT;
    ptx::mbar_arrive_expect_tx(&bar[sl], (uint32_t)(SLOT * sizeof(bf16)));
    ptx::tma_load_4d(dst, &tm_g, &bar[sl], c0, w0 - 1, r0 - 1, b0);
    ptx::tma_load_4d(dst + G_ELEMS, &tm_e, &bar[sl], c0, w0, r0, b0);
  };
  int t = worker;
  if (tid == 0 && t < g.ntiles) issue(t, 0);
  for (int it = 0; t < g.ntiles; t += g.nworkers, ++it) {
    int b0, r0, w0;
    tile_coords(t, b0, r0, w0);
    const int sl = it & 1;
    if (tid == 0 && t + g.nworkers < g.ntiles) issue(t + g.nworkers, sl ^ 1);
    ptx::mbar_wait(&bar[sl], (it >> 1) & 1);
    const int b = b0 + img;
    if (b < g.B) {
      const bf16* gcol = slot0 + sl * SLOT + ((img * TR2 + rs) * W2 + x0) * DWB_CC + cg * 2;
      const bf16* ecol = slot0 + sl * SLOT + G_ELEMS + ((img * TRT + rs) * W_ + x0) * DWB_CC + cg * 2;
      const int h0 = r0 + rs;
      bf16* out = du1 + (((long long)b * g.H + h0) * Wi + w0 + x0) * g.Cm + c;
      const long long row_stride = (long long)Wi * g.Cm;
      f32x2 win[3][4];
      auto load_row = [&](int j, f32x2 (&row)[4]) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(gcol + j * GROW + cc * DWB_CC);
          row[cc] = pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
        }
      };
      load_row(0, win[0]);
      load_row(1, win[1]);
#pragma unroll
      for (int j = 0; j < TRS; ++j) {
        load_row(j + 2, win[(j + 2) % 3]);
        if (h0 + j < g.H) {  // warp-uniform: the lanes of a warp share (image, row)
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const uint32_t eu = *reinterpret_cast<const uint32_t*>(ecol + j * EROW + p * DWB_CC);
            const f32x2 ev = pk2(__uint_as_float(eu << 16), __uint_as_float(eu & 0xffff0000u));
            float u0, u1, a0, a1, d0, d1;
            unpk2(fma2(ev, sc2, sh2), u0, u1);
            act_both_t<ACT, true>(u0, &a0, &d0);
            act_both_t<ACT, true>(u1, &a1, &d1);
            const f32x2 ea = pk2(a0, a1);
            f32x2 de = 0ull;
#pragma unroll
            for (int s9 = 0; s9 < 9; ++s9) {
              const f32x2 gv = win[(j + s9 / 3) % 3][s9 % 3 + p];
              de = fma2(wf[s9], gv, de);
              dwa[s9] = fma2(ea, gv, dwa[s9]);
            }
            const f32x2 o = mul2(de, pk2(d0, d1));
            float o0, o1;
            unpk2(o, o0, o1);
            const __nv_bfloat162 ob = __floats2bfloat162_rn(o0, o1);
            *reinterpret_cast<uint32_t*>(out + j * row_stride + p * g.Cm) = *reinterpret_cast<const uint32_t*>(&ob);
            a_db = add2(a_db, o);
            a_dg = fma2(o, ev, a_dg);  // sum o*e; xhat = (e - mean)*rstd is folded in at the end
          }
        }
      }
    }
    __syncthreads();  // slot `sl` is free for the TMA issued at the top of the next iteration
  }
  // fold: lanes (cg, cg + 16) of a warp share the channel pair; then shared-memory atomics, one global flush per CTA
  {
    float v[22];
#pragma unroll
    for (int s9 = 0; s9 < 9; ++s9) unpk2(dwa[s9], v[2 * s9], v[2 * s9 + 1]);
    unpk2(a_db, v[18], v[19]);
    unpk2(a_dg, v[20], v[21]);
#pragma unroll
    for (int i = 0; i < 22; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
    if ((tid & 16) == 0) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int ch = cg * 2 + k;
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) atomicAdd(&s_dw[ch * 9 + 8 - s9], v[2 * s9 + k]);
        atomicAdd(&s_db[ch], v[18 + k]);
        atomicAdd(&s_dg[ch], rstd1[c0 + ch] * (v[20 + k] - mean1[c0 + ch] * v[18 + k]));
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < DWB_CC * 9; i += DW_THREADS) atomicAdd(dwgt + (long long)c0 * 9 + i, s_dw[i]);
  if (tid < DWB_CC) {
    atomicAdd(dbeta1 + c0 + tid, s_db[tid]);
    atomicAdd(dgamma1 + c0 + tid, s_dg[tid]);
  }
}

template <int ACT, int W_>
int dwb_walk_launch(const void* dd_pre, const void* e_pre, const float* scale1, const float* shift1, const float* mean1,
                    const float* rstd1, const float* w, void* du1, float* dw, float* dgamma1, float* dbeta1, int B, int H,
                    int W, int Cm, cudaStream_t st) {
  using S = DwwShape<W_>;
  constexpr int TRT = S::TRS * S::NSB;
  constexpr int SLOT_BYTES = (S::NI * (TRT + 2) * (W_ + 2) + S::NI * TRT * W_) * DWB_CC * 2;
  const int smem = 2 * SLOT_BYTES;
  auto kern = dwconv_bwd_walk_kernel<ACT, W_>;
  int occ = 1;
  if (int rc = dw_smem_optin(kern, smem, &occ)) return rc;
  DwwGeom g;
  g.B = B; g.H = H; g.W = W; g.Cm = Cm;
  g.bands = (H + TRT - 1) / TRT;
  g.tiles_w = W / W_;
  const long long nt = (long long)((B + S::NI - 1) / S::NI) * g.bands * g.tiles_w;
  if (nt > 0x7fffffffLL) { ogv_set_error("dwconv_bwd: too many tiles"); return OGV_ERR_UNSUPPORTED; }
  g.ntiles = (int)nt;
  g.nchunks = Cm / DWB_CC;
  long long want = ((long long)ogv_num_sms() * occ) / g.nchunks;
  if (want > nt) want = nt;
  if (want < 1) want = 1;
  g.nworkers = (int)want;
  CUtensorMap tmg, tme;
  {
    unsigned long long dims[4] = {(unsigned long long)Cm, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
    unsigned long long str[3] = {(unsigned long long)Cm * 2, (unsigned long long)W * Cm * 2, (unsigned long long)H * W * Cm * 2};
    unsigned boxg[4] = {(unsigned)DWB_CC, (unsigned)(W_ + 2), (unsigned)(TRT + 2), (unsigned)S::NI};
    unsigned boxe[4] = {(unsigned)DWB_CC, (unsigned)W_, (unsigned)TRT, (unsigned)S::NI};
    if (int rc = ogv_make_tmap(&tmg, dd_pre, OGV_BF16, 4, dims, str, boxg, 0)) return rc;
    if (int rc = ogv_make_tmap(&tme, e_pre, OGV_BF16, 4, dims, str, boxe, 0)) return rc;
  }
  kern<<<g.nchunks * g.nworkers, DW_THREADS, smem, st>>>(tmg, tme, scale1, shift1, mean1, rstd1, w,
                                                         reinterpret_cast<bf16*>(du1), dw, dgamma1, dbeta1, g);
  return ogv_check_launch("dwconv_bwd");
}

}  // namespace

extern "C" int ogv_dwconv_fwd(const void* e_pre, const float* scale1, const float* shift1, const float* w,
                              void* d_pre, float* sum2, float* sumsq2, int B, int H, int W, int Cm, int act,
                              int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(e_pre && scale1 && shift1 && w && d_pre, "dwconv_fwd: null pointer");
  OGV_REQUIRE(Cm > 0 && Cm % 8 == 0 && H > 0 && W > 0, "dwconv_fwd: channels must be a multiple of 8");
  OGV_REQUIRE((reinterpret_cast<uintptr_t>(e_pre) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_pre) & 15) == 0,
              "dwconv_fwd: tensors must be 16-byte aligned");
  static int variant = -1;  // OGV_DW_FWD=tile | sweep select the older kernels (A/B measurements); default: prefetched sweep
  if (variant < 0) {
    const char* e = getenv("OGV_DW_FWD");
    // measured (stage shapes of cfg 2): the ring does NOT help -- 549 vs 519 us at stage 0: the sweep is bound by its
    // ~30 issued instructions per output (halo rows make it activate every input twice), not by load latency.  The
    // plain sweep stays the default; OGV_DW_FWD=pf / tile select the other two for A/B runs.
    variant = (e && e[0] == 't') ? 1 : ((e && e[0] == 'p') ? 2 : ((e && e[0] == 's') ? 0 : 3));
  }
  if (variant == 3 && dtype == OGV_BF16 && (W == 4 || W == 8 || W == 16 || (W >= 32 && W % 32 == 0)) && Cm % DWW_CC == 0 &&
      (reinterpret_cast<uintptr_t>(scale1) & 15) == 0 && (reinterpret_cast<uintptr_t>(shift1) & 15) == 0) {
    cudaStream_t st = (cudaStream_t)stream;
    OGV_DISPATCH_ACT(act, ACT, {
      switch (W) {
        case 16: return dww_launch<ACT, 16>(e_pre, scale1, shift1, w, d_pre, sum2, sumsq2, B, H, W, Cm, st);
        case 8: return dww_launch<ACT, 8>(e_pre, scale1, shift1, w, d_pre, sum2, sumsq2, B, H, W, Cm, st);
        case 4: return dww_launch<ACT, 4>(e_pre, scale1, shift1, w, d_pre, sum2, sumsq2, B, H, W, Cm, st);
        default: return dww_launch<ACT, 32>(e_pre, scale1, shift1, w, d_pre, sum2, sumsq2, B, H, W, Cm, st);
      }
    });
  }
  if (variant == 2) {
    constexpr int RR = 2, KD = 6;
    const int bands = (H + RR - 1) / RR;
    const long long nb = (long long)B * bands;
    OGV_REQUIRE(nb < 0x7fffffffLL, "dwconv_fwd: too many row bands");
    const int ychunks = ogv_ceil_div(Cm, 128);
    long long gx = (nb + 7) / 8;
    // persistent: exactly the resident CTAs (2 per SM), every warp walks its bands with the ring running ahead
    const long long cap = ((long long)ogv_num_sms() * 2 + ychunks - 1) / ychunks;
    if (gx > cap) gx = cap;
    OGV_DISPATCH_DTYPE(dtype, T, {
      constexpr int smem = 8 * KD * (RR + 2) * 32 * 4 * (int)sizeof(T);
      OGV_DISPATCH_ACT(act, ACT, {
        static bool attr = false;
        if (!attr) {
          cudaFuncSetAttribute(dwconv_fwd_sweep_pf_kernel<T, ACT, RR, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          attr = true;
        }
        dwconv_fwd_sweep_pf_kernel<T, ACT, RR, KD><<<dim3((unsigned)gx, ychunks), 256, smem, (cudaStream_t)stream>>>(
            reinterpret_cast<const T*>(e_pre), scale1, shift1, w, reinterpret_cast<T*>(d_pre), sum2, sumsq2, B, H, W,
            Cm, bands, (int)nb);
      });
      return ogv_check_launch("dwconv_fwd");
    });
  }
  if (variant == 0 || variant == 3) {
    constexpr int RR = 2;
    const int bands = (H + RR - 1) / RR;
    const long long nb = (long long)B * bands;
    OGV_REQUIRE(nb < 0x7fffffffLL, "dwconv_fwd: too many row bands");
    const int ychunks = ogv_ceil_div(Cm, 128);
    long long gx = (nb + 7) / 8;
    const long long cap = ((long long)ogv_num_sms() * 2 + ychunks - 1) / ychunks * 4;  // a few waves of 2 CTAs/SM
    if (gx > cap) gx = cap;
    OGV_DISPATCH_DTYPE(dtype, T, {
      OGV_DISPATCH_ACT(act, ACT, {
        dwconv_fwd_sweep_kernel<T, ACT, RR><<<dim3((unsigned)gx, ychunks), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const T*>(e_pre), scale1, shift1, w, reinterpret_cast<T*>(d_pre), sum2, sumsq2, B, H, W,
            Cm, bands, (int)nb);
      });
      return ogv_check_launch("dwconv_fwd");
    });
  }
  OGV_DISPATCH_DTYPE(dtype, T, {
    constexpr int CC = dw_cc<T>();
    const int smem = DW_BUF_POS * CC * (int)(sizeof(T) + sizeof(float)) + 64;
    OGV_DISPATCH_ACT(act, ACT, {
      int occ = 1;
      if (int rc = dw_smem_optin(dwconv_fwd_kernel<T, CC, ACT>, smem, &occ)) return rc;
      DwGeom g;
      if (dw_make_geom(B, H, W, Cm, CC, occ, &g)) { ogv_set_error("dwconv_fwd: cannot tile %dx%d", H, W); return OGV_ERR_UNSUPPORTED; }
      CUtensorMap tm;
      if (int rc = dw_tmap<T>(&tm, e_pre, dtype, g, CC, 1)) return rc;
      dwconv_fwd_kernel<T, CC, ACT><<<g.nchunks * g.nworkers, DW_THREADS, smem, (cudaStream_t)stream>>>(
          tm, scale1, shift1, w, reinterpret_cast<T*>(d_pre), sum2, sumsq2, g);
    });
    return ogv_check_launch("dwconv_fwd");
  });
}

#define OGV_DW_BWD_SPEC(TWc, THc, NIc)                                                                              \
  if (full && g.TW == TWc && g.TH == THc && g.NI == NIc) {                                                          \
auto kern = dwconv_bwd_kernel<T, CC, ACT, TWc, THc, NIc>;                                                       \
int occ2 = 1;                                                                                                   \
if (int rc = dw_smem_optin(kern, smem, &occ2)) return rc;                                                       \
if (occ2 >= occ) {                                                                                              \
  kern<<<g.nchunks * g.nworkers, DW_THREADS, smem, (cudaStream_t)stream>>>(                                     \
      tmg, tme, scale1, shift1, mean1, rstd1, w, reinterpret_cast<T*>(du1), dw, dgamma1, dbeta1, g);            \
  return ogv_check_launch("dwconv_bwd");                                                                        \
}                                                                                                               \
  }

extern "C" int ogv_dwconv_bwd(const void* dd_pre, const void* e_pre, const float* scale1, const float* shift1,
                              const float* mean1, const float* rstd1, const float* w, void* du1, float* dw,
                              float* dgamma1, float* dbeta1, int B, int H, int W, int Cm, int act, int dtype,
                              void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(dd_pre && e_pre && scale1 && shift1 && mean1 && rstd1 && w && du1 && dw && dgamma1 && dbeta1,
              "dwconv_bwd: null pointer");
  OGV_REQUIRE(Cm > 0 && Cm % 8 == 0 && H > 0 && W > 0, "dwconv_bwd: channels must be a multiple of 8");
  OGV_REQUIRE((reinterpret_cast<uintptr_t>(dd_pre) & 15) == 0 && (reinterpret_cast<uintptr_t>(e_pre) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(du1) & 15) == 0,
              "dwconv_bwd: tensors must be 16-byte aligned");
  static int walk = -1;  // OGV_DW_BWD=tile selects the older TMA-tiled strip kernel (A/B measurements)
  if (walk < 0) { const char* e = getenv("OGV_DW_BWD"); walk = (e && e[0] == 't') ? 0 : 1; }
  if (walk && dtype == OGV_BF16 && (W == 4 || W == 8 || W == 16 || (W >= 32 && W % 32 == 0)) && Cm % DWB_CC == 0) {
    cudaStream_t st = (cudaStream_t)stream;
    OGV_DISPATCH_ACT(act, ACT, {
      switch (W) {
        case 16: return dwb_walk_launch<ACT, 16>(dd_pre, e_pre, scale1, shift1, mean1, rstd1, w, du1, dw, dgamma1, dbeta1, B, H, W, Cm, st);
        case 8: return dwb_walk_launch<ACT, 8>(dd_pre, e_pre, scale1, shift1, mean1, rstd1, w, du1, dw, dgamma1, dbeta1, B, H, W, Cm, st);
        case 4: return dwb_walk_launch<ACT, 4>(dd_pre, e_pre, scale1, shift1, mean1, rstd1, w, du1, dw, dgamma1, dbeta1, B, H, W, Cm, st);
        default: return dwb_walk_launch<ACT, 32>(dd_pre, e_pre, scale1, shift1, mean1, rstd1, w, du1, dw, dgamma1, dbeta1, B, H, W, Cm, st);
      }
    });
  }
  OGV_DISPATCH_DTYPE(dtype, T, {
    constexpr int CC = dw_cc<T>();
    const int smem = DW_BUF_POS * CC * (int)(sizeof(T) + sizeof(float)) + 2 * 256 * CC * (int)sizeof(T) + 64;
    OGV_DISPATCH_ACT(act, ACT, {
      int occ = 1;
      if (int rc = dw_smem_optin(dwconv_bwd_kernel<T, CC, ACT, 0, 0, 0>, smem, &occ)) return rc;
      DwGeom g;
      if (dw_make_geom(B, H, W, Cm, CC, occ, &g)) { ogv_set_error("dwconv_bwd: cannot tile %dx%d", H, W); return OGV_ERR_UNSUPPORTED; }
      CUtensorMap tmg, tme;
      if (int rc = dw_tmap<T>(&tmg, dd_pre, dtype, g, CC, 1)) return rc;
      if (int rc = dw_tmap<T>(&tme, e_pre, dtype, g, CC, 0)) return rc;
      // the 32- and 16-pixel square-image geometries (stages 0-1 of the 32 px configs, 1-2 of the 64 px ones) are compiled
      // with constant tile shapes; anything else (ragged edges, other sizes) runs the generic instantiation.
      // OGV_DW_BWD_GENERIC=1: A/B
      static int generic = -1;
      if (generic < 0) { const char* e = getenv("OGV_DW_BWD_GENERIC"); generic = (e && e[0] == '1') ? 1 : 0; }
      const bool full = !generic && W % g.TW == 0 && H % g.TH == 0 && g.TH % 2 == 0;
      OGV_DW_BWD_SPEC(32, 8, 1)   // 745 -> 610 us at stage 0 of cfg 2
      OGV_DW_BWD_SPEC(16, 16, 1)  // 394 -> 387 us at stage 1
      // (8, 8, 3) and (4, 4, 9) measured 2 % / 10 % SLOWER than the generic kernel (the folded offsets let ptxas hoist all
      // twelve tile loads of an item, 100 bytes of spills at the 128-register cap): not instantiated
      dwconv_bwd_kernel<T, CC, ACT, 0, 0, 0><<<g.nchunks * g.nworkers, DW_THREADS, smem, (cudaStream_t)stream>>>(
          tmg, tme, scale1, shift1, mean1, rstd1, w, reinterpret_cast<T*>(du1), dw, dgamma1, dbeta1, g);
    });
    return ogv_check_launch("dwconv_bwd");
  });
}
