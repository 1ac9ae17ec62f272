// Outlook attention core: softmax over the k*k=9 logits of each head, then the weighted gather over
// the 3x3 neighbourhood of v (zero outside the image, weights NOT renormalised) -- the reference's
// F.unfold + mul + sum (outlook_attention.py:104-120) without ever materialising the 9x im2col tensor.
// Backward is the overlap-add (col2im) written as a gather plus the softmax Jacobian.
//
// `va` rows: [ v (C) | logits head0 (9) | logits head1 (9) | ... | zero padding ] with stride ld_va.
//
// A CTA owns a band of TR full-width rows of one image.  The softmax is evaluated ONCE per
// (position, head) into shared memory (backward: also for the halo rows, whose probabilities the
// overlap-add needs); the channel work then runs with one thread per (position, VEC channels): 16-byte
// neighbour loads that hit L1 (each row of the band is touched by 9 taps of the same CTA), the 9
// probabilities as shared-memory broadcasts.  No integer division on the data path.
#include <stdlib.h>

#include "ogv_common.cuh"
#include "ogv_ptx.cuh"
#include "ogv_tma.cuh"  // FastDiv, tensor maps
#include "../../include/ogv.h"

namespace {

constexpr int OL_THREADS = 256;

struct OlGeom {
  int B, H, W, C, heads, hd;
  int TR, bands;          // rows per band, bands per image
  long long ld;           // row stride of va / dva
  FastDiv d_w, d_heads, d_hd, d_nvc;  // x / W, x / heads, x / head_dim, x / (C / VEC) for x < 65536
};

__device__ __forceinline__ void softmax9(const float (&l)[9], float (&a)[9]) {
  float mx = l[0];
#pragma unroll
  for (int t = 1; t < 9; ++t) mx = fmaxf(mx, l[t]);
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    a[t] = __expf(l[t] - mx);
    s += a[t];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int t = 0; t < 9; ++t) a[t] *= inv;
}

// probabilities of rows [row_lo, row_hi) of image b -> A_s[((row - row_lo) * W + w) * heads + head][9]
template <typename T>
__device__ __forceinline__ void ol_softmax_rows(const T* __restrict__ va, float* __restrict__ A_s, const OlGeom& g,
                                                int b, int row_lo, int row_hi) {
  const int n = (row_hi - row_lo) * g.W * g.heads;
  const T* base = va + ((long long)b * g.H + row_lo) * g.W * g.ld + g.C;
  for (int i = threadIdx.x; i < n; i += OL_THREADS) {
    const int pos = fdiv(i, g.d_heads);
    const int head = i - pos * g.heads;
    const T* lp = base + (long long)pos * g.ld + head * 9;
    float l[9], a[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) l[t] = ld1(lp + t);
    softmax9(l, a);
#pragma unroll
    for (int t = 0; t < 9; ++t) A_s[i * 9 + t] = a[t];
  }
}

// ------------------------------------------------------------------------------------------------
// forward: y[p,c] = sum_t A[p, head(c), t] * v[p + d_t, c]
// All 9 neighbour loads are unconditional (out-of-image taps read the centre row and get weight 0), so
// they are issued back to back instead of one dependent load per branch.
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(OL_THREADS) outlook_fwd_kernel(const T* __restrict__ va, T* __restrict__ y,
                                                                const OlGeom g) {
  extern __shared__ __align__(16) float A_s[];  // [TR*W*heads][9]
  const int b = blockIdx.x / g.bands;
  const int r0 = (blockIdx.x - b * g.bands) * g.TR;
  const int r1 = min(g.H, r0 + g.TR);
  ol_softmax_rows(va, A_s, g, b, r0, r1);
  __syncthreads();
  const int nvc = g.C / VEC;
  const int npos = (r1 - r0) * g.W;
  const long long img = (long long)b * g.H * g.W;
  for (int i = threadIdx.x; i < npos * nvc; i += OL_THREADS) {
    const int pos = fdiv(i, g.d_nvc);
    const int c = (i - pos * nvc) * VEC;
    const int hrow = fdiv(pos, g.d_w);
    const int w = pos - hrow * g.W;
    const int h = r0 + hrow;
    const float* a = A_s + (pos * g.heads + fdiv(c, g.d_hd)) * 9;
    const long long m = img + (long long)h * g.W + w;
    const T* vc = va + m * g.ld + c;
    const bool rok[3] = {h >= 1, true, h + 1 < g.H};
    const bool cok[3] = {w >= 1, true, w + 1 < g.W};
    const int ldv_ = (int)g.ld, rowv = g.W * ldv_;
    float v[9][VEC], wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int dh = t / 3 - 1, dw = t % 3 - 1;
      const bool ok = rok[dh + 1] && cok[dw + 1];
      ldv<VEC>(vc + (ok ? dh * rowv + dw * ldv_ : 0), v[t]);
      wt[t] = ok ? a[t] : 0.f;
    }
    float acc[VEC];
    if constexpr (VEC % 2 == 0) {
      f32x2 acc2[VEC / 2];
#pragma unroll
      for (int k = 0; k < VEC / 2; ++k) acc2[k] = 0ull;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const f32x2 w2 = pk2(wt[t], wt[t]);
#pragma unroll
        for (int k = 0; k < VEC / 2; ++k) acc2[k] = fma2(w2, pk2(v[t][2 * k], v[t][2 * k + 1]), acc2[k]);
      }
#pragma unroll
      for (int k = 0; k < VEC / 2; ++k) unpk2(acc2[k], acc[2 * k], acc[2 * k + 1]);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = fmaf(wt[t], v[t][k], acc[k]);
    }
    stv<VEC>(y + m * g.C + c, acc);
  }
}

// zero the alignment padding [from, to) of one dva row (a few elements): 4-byte stores where the layout allows
template <typename T>
__device__ __forceinline__ void ol_zero_pad(T* row, int from, int to) {
  if (sizeof(T) == 2 && ((from | to) & 1) == 0 && (reinterpret_cast<uintptr_t>(row) & 3) == 0) {
    for (int j = from; j < to; j += 2) *reinterpret_cast<uint32_t*>(row + j) = 0u;
  } else {
    for (int j = from; j < to; ++j) st1(row + j, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// backward:  dv[q,c]     = sum_t A[head(c), t, q - d_t] * dy[q - d_t, c]           (overlap-add as a gather)
//            dA[p,h,t]   = sum_{c in head h} dy[p,c] * v[p + d_t, c]
//            dlogits     = A * (dA - sum_t A * dA)
// dva row: [ dv (C) | dlogits (9*heads) | zeros up to ld ]
// SHFL: the hd/VEC threads of a (position, head) are an aligned lane group -> dA is reduced with shuffles
// and the group leader finishes the softmax Jacobian; otherwise shared-memory atomics + a second phase.
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, bool SHFL>
__global__ void __launch_bounds__(OL_THREADS) outlook_bwd_kernel(const T* __restrict__ va, const T* __restrict__ dy,
                                                                T* __restrict__ dva, const OlGeom g) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / g.bands;
  const int r0 = (blockIdx.x - b * g.bands) * g.TR;
  const int r1 = min(g.H, r0 + g.TR);
  const int h_lo = max(0, r0 - 1), h_hi = min(g.H, r1 + 1);  // rows whose probabilities are needed
  float* A_s = sm;                                            // [(h_hi-h_lo)*W*heads][9]
  float* dA_s = sm + (g.TR + 2) * g.W * g.heads * 9;          // [TR*W*heads][9]  (atomics path only)
  const int npos = (r1 - r0) * g.W;
  if (!SHFL)
    for (int i = threadIdx.x; i < npos * g.heads * 9; i += OL_THREADS) dA_s[i] = 0.f;
  ol_softmax_rows(va, A_s, g, b, h_lo, h_hi);
  __syncthreads();

  const int nvc = g.C / VEC;
  const int ntot = npos * nvc;
  const int gs = g.hd / VEC;  // threads per (position, head)
  const int nl = g.heads * 9;
  const long long img = (long long)b * g.H * g.W;
  for (int i0 = 0; i0 < ntot; i0 += OL_THREADS) {
    const int iraw = i0 + threadIdx.x;
    const bool valid = iraw < ntot;
    const int i = valid ? iraw : ntot - 1;  // idle lanes shadow the last item (they still take part in shuffles)
    const int pos = fdiv(i, g.d_nvc);
    const int c = (i - pos * nvc) * VEC;
    const int head = fdiv(c, g.d_hd);
    const int hrow = fdiv(pos, g.d_w);
    const int w = pos - hrow * g.W;
    const int h = r0 + hrow;
    const long long m = img + (long long)h * g.W + w;
    const T* vc = va + m * g.ld + c;
    const T* gc = dy + m * g.C + c;
    float gctr[VEC], dv[VEC], dA[9];
    ldv<VEC>(gc, gctr);
    // tap validity as 3 row x 3 column flags and 32-bit element offsets (the per-tap 64-bit index arithmetic and
    // the twice-evaluated bounds checks were over half of this kernel's instructions)
    const bool rok[3] = {h >= 1, true, h + 1 < g.H};
    const bool cok[3] = {w >= 1, true, w + 1 < g.W};
    const int ldv_ = (int)g.ld, rowv = g.W * ldv_, rowg = g.W * g.C;
    const int hw9 = g.heads * 9, whw9 = g.W * hw9;
    const int abase = (((h - h_lo) * g.W + w) * g.heads + head) * 9;
    // two batches of unconditional loads: v at p + d_t (for dA), dy at q - d_t (for dv)
    {
      float v[9][VEC];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dh = t / 3 - 1, dw = t % 3 - 1;
        const bool ok = rok[dh + 1] && cok[dw + 1];
        ldv<VEC>(vc + (ok ? dh * rowv + dw * ldv_ : 0), v[t]);
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dh = t / 3 - 1, dw = t % 3 - 1;
        const bool ok = rok[dh + 1] && cok[dw + 1];
        float sdot = 0.f;
        if constexpr (VEC % 2 == 0) {
          f32x2 acc2 = 0ull;
#pragma unroll
          for (int k = 0; k < VEC; k += 2) acc2 = fma2(pk2(gctr[k], gctr[k + 1]), pk2(v[t][k], v[t][k + 1]), acc2);
          float lo, hi;
          unpk2(acc2, lo, hi);
          sdot = lo + hi;
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) sdot = fmaf(gctr[k], v[t][k], sdot);
        }
        dA[t] = ok ? sdot : 0.f;
      }
    }
    {
      float gs_[9][VEC], at[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int dh = t / 3 - 1, dw = t % 3 - 1;
        // source position p = q - d_t used tap t to read q
        const bool ok = rok[1 - dh] && cok[1 - dw];
        ldv<VEC>(gc - (ok ? dh * rowg + dw * g.C : 0), gs_[t]);
        const float a = A_s[ok ? abase - dh * whw9 - dw * hw9 + t : 0];
        at[t] = ok ? a : 0.f;
      }
      if constexpr (VEC % 2 == 0) {
        f32x2 dv2[VEC / 2];
#pragma unroll
        for (int k = 0; k < VEC / 2; ++k) dv2[k] = 0ull;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const f32x2 a2 = pk2(at[t], at[t]);
#pragma unroll
          for (int k = 0; k < VEC / 2; ++k) dv2[k] = fma2(a2, pk2(gs_[t][2 * k], gs_[t][2 * k + 1]), dv2[k]);
        }
#pragma unroll
        for (int k = 0; k < VEC / 2; ++k) unpk2(dv2[k], dv[2 * k], dv[2 * k + 1]);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) dv[k] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int k = 0; k < VEC; ++k) dv[k] = fmaf(at[t], gs_[t][k], dv[k]);
      }
    }
    if (valid) stv<VEC>(dva + m * g.ld + c, dv);
    if (SHFL) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
        for (int off = gs >> 1; off > 0; off >>= 1) dA[t] += __shfl_xor_sync(0xffffffffu, dA[t], off);
      if (valid && (c - head * g.hd) == 0) {  // group leader: softmax Jacobian for this (position, head)
        const float* a = A_s + ((((r0 - h_lo) + hrow) * g.W + w) * g.heads + head) * 9;
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) dot = fmaf(a[t], dA[t], dot);
        T* out = dva + m * g.ld + g.C + head * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) st1(out + t, a[t] * (dA[t] - dot));
        if (head == 0) ol_zero_pad(dva + m * g.ld, g.C + nl, (int)g.ld);
      }
    } else if (valid) {
      float* dst = dA_s + (pos * g.heads + head) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dst + t, dA[t]);
    }
  }
  if (SHFL) return;
  __syncthreads();
  // softmax Jacobian, one thread per (position, head); also zero the alignment padding of the row
  for (int i = threadIdx.x; i < npos * g.heads; i += OL_THREADS) {
    const int pos = fdiv(i, g.d_heads);
    const int head = i - pos * g.heads;
    const int hrow = fdiv(pos, g.d_w);
    const float* a = A_s + ((((r0 - h_lo) + hrow) * g.W + (pos - hrow * g.W)) * g.heads + head) * 9;
    const float* d = dA_s + i * 9;
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) dot = fmaf(a[t], d[t], dot);
    T* out = dva + (img + (long long)r0 * g.W + pos) * g.ld + g.C + head * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) st1(out + t, a[t] * (d[t] - dot));
    if (head == 0) ol_zero_pad(dva + (img + (long long)r0 * g.W + pos) * g.ld, g.C + nl, (int)g.ld);
  }
}

inline int pick_vec(int C, int hd, long long ld) {
  if (hd % 8 == 0 && C % 8 == 0 && ld % 8 == 0) return 8;
  if (hd % 4 == 0 && C % 4 == 0 && ld % 4 == 0) return 4;
  return 1;
}

int ol_geom(int B, int H, int W, int C, int heads, long long ld, int vec, OlGeom* g) {
  g->B = B; g->H = H; g->W = W; g->C = C; g->heads = heads; g->hd = C / heads; g->ld = ld;
  int tr = 128 / W;
  if (tr < 1) tr = 1;
  if (tr > H) tr = H;
  // keep the probability tables within ~40 KB of shared memory
  while (tr > 1 && (long long)(2 * tr + 2) * W * heads * 9 * 4 > 40 * 1024) --tr;
  g->TR = tr;
  g->bands = (H + tr - 1) / tr;
  if ((long long)(tr + 2) * W * heads >= 65536 || (long long)(2 * tr + 2) * W * heads * 9 * 4 > 200 * 1024) return -1;
  g->d_w = make_fastdiv(W);
  g->d_heads = make_fastdiv(heads);
  g->d_hd = make_fastdiv(g->hd);
  g->d_nvc = make_fastdiv(C / vec);
  if (C >= 65536 || (long long)tr * W * (C / vec) >= 65536) return -1;
  return 0;
}

template <typename K>
int ol_optin(K kernel, size_t bytes) {
  if (bytes > 40 * 1024) {  // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      ogv_set_error("outlook: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return OGV_ERR_CUDA;
    }
  }
  return OGV_OK;
}

// ================================================================================================
// Tiled kernels (bf16, W in {4, 8, 16, 32}, C % 64 == 0): the hot configuration of every model stage.
// ncu on the kernels above (stage 0 of cfg 2): 49 (forward) / 108 (backward) issued instructions per output element for
// 4.5 / 9 FFMA2 -- every tap re-loads and re-converts its bf16 neighbour and pays 64-bit address arithmetic plus validity
// selects.  Here a CTA owns (NI images x TRT rows x 64 channels): ONE TMA box (cp.async.bulk.tensor.4d over
// [ld, W, H, B], borders and ragged edges zero-filled by the descriptor) lands the halo tile in shared memory while the
// CTA evaluates the softmax of its positions; then every thread is a WALKER: it owns two adjacent columns x four
// channels and moves down the rows of its sub-band with the 3 x 4 (rows x columns) fp32 window in registers -- one
// new row (4 LDS.64 + 16 conversions) per 8 outputs, no predicates (the tile has the zero halo), no index decode.
// ================================================================================================
constexpr int OLT_CC = 64;
struct OltGeom {
  int B, H, W, C, heads, hd, hpc;  // W = IMAGE width (a multiple of the tile width); hpc = heads per 64-channel chunk
  long long ld;
  int TRT, NI, NSB, bands, nchunks, tiles_w;
};
template <int W_> struct OltShape;
template <> struct OltShape<32> { static constexpr int NI = 1, NSB = 1, TRS = 8; };
template <> struct OltShape<16> { static constexpr int NI = 1, NSB = 2, TRS = 8; };
template <> struct OltShape<8> { static constexpr int NI = 4, NSB = 1, TRS = 8; };
template <> struct OltShape<4> { static constexpr int NI = 8, NSB = 1, TRS = 4; };

// one tile row (4 columns x 4 bf16 channels) -> fp32 channel pairs
__device__ __forceinline__ void olt_load_row(const bf16* p, f32x2 (&row)[4][2]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint2 u = *reinterpret_cast<const uint2*>(p + c * OLT_CC);
    row[c][0] = pk2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
    row[c][1] = pk2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
  }
}
__device__ __forceinline__ uint32_t olt_pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// The walker: out[p][ch] = sum_s wt(j, p)[s] * tile[(row rs + j - 1 + s / 3), (col x0 + p - 1 + s % 3)][ch] for the two
// columns (x0, x0 + 1) of TRS consecutive rows.  `tile_col0` points at tile row rs (halo coordinates), tile column x0,
// channel cg*4; `getw(j, p, av)` delivers the nine weights of output (row rs + j, column x0 + p).
template <int W_, int TRS, typename GetW, typename Store>
__device__ __forceinline__ void olt_walk(const bf16* tile_col0, GetW&& getw, Store&& store) {
  constexpr int ROW = (W_ + 2) * OLT_CC;
  f32x2 win[3][4][2];
  olt_load_row(tile_col0, win[0]);
  olt_load_row(tile_col0 + ROW, win[1]);
#pragma unroll
  for (int j = 0; j < TRS; ++j) {
    olt_load_row(tile_col0 + (j + 2) * ROW, win[(j + 2) % 3]);
    f32x2 acc[2][2] = {{0ull, 0ull}, {0ull, 0ull}};
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      float av[9];
      getw(j, p, av);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const f32x2 w2 = pk2(av[t], av[t]);
        const int r = (j + t / 3) % 3, c = t % 3 + p;
        acc[p][0] = fma2(w2, win[r][c][0], acc[p][0]);
        acc[p][1] = fma2(w2, win[r][c][1], acc[p][1]);
      }
    }
    store(j, acc);
  }
}

template <int W_>
__global__ void __launch_bounds__(OL_THREADS) outlook_fwd_tile_kernel(const __grid_constant__ CUtensorMap tm_v,
                                                                     const bf16* __restrict__ va, bf16* __restrict__ y,
                                                                     const OltGeom g) {
  using S = OltShape<W_>;
  constexpr int NI = S::NI, NSB = S::NSB, TRS = S::TRS, TRT = TRS * NSB;
  constexpr int TILE_ELEMS = NI * (TRT + 2) * (W_ + 2) * OLT_CC;
  extern __shared__ __align__(128) uint8_t olt_sm[];
  bf16* const tile = reinterpret_cast<bf16*>(olt_sm);
  float* const A_s = reinterpret_cast<float*>(olt_sm + TILE_ELEMS * sizeof(bf16));  // [NI*TRT*W][hpc][12]
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % g.nchunks;
  int rest = blockIdx.x / g.nchunks;
  const int w0 = (rest % g.tiles_w) * W_;  // images wider than the tile (W = 64, 96, ...) are cut into column tiles
  rest /= g.tiles_w;
  const int band = rest % g.bands;
  const int b0 = (rest / g.bands) * NI, r0 = band * TRT, c0 = chunk * OLT_CC;
  const int head0 = c0 / g.hd, hpc = g.hpc;
  const int Wi = g.W;
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(&bar, (uint32_t)(TILE_ELEMS * sizeof(bf16)));
    ptx::tma_load_4d(tile, &tm_v, &bar, c0, w0 - 1, r0 - 1, b0);
  }
  // softmax of the tile's centre positions while the box is in flight
  const int nsm = NI * TRT * W_ * hpc;
  for (int i = tid; i < nsm; i += OL_THREADS) {
    const int hl = i % hpc, pos = i / hpc;
    const int w = pos % W_, rr = (pos / W_) % TRT, img = pos / (W_ * TRT);
    const int b = b0 + img, h = r0 + rr;
    float a[9];
    if (b < g.B && h < g.H) {
      const bf16* lp = va + (((long long)b * g.H + h) * Wi + w0 + w) * g.ld + g.C + (head0 + hl) * 9;
      float l[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) l[t] = ld1(lp + t);
      softmax9(l, a);
    } else {
#pragma unroll
      for (int t = 0; t < 9; ++t) a[t] = 0.f;
    }
    float* dst = A_s + i * 12;
    *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(a[4], a[5], a[6], a[7]);
    dst[8] = a[8];
  }
  __syncthreads();  // A_s complete, barrier initialised for everyone
  ptx::mbar_wait(&bar, 0);

  const int cg = tid % 16;
  const int xp = (tid / 16) % (W_ / 2);
  const int grp = tid / (8 * W_);
  const int img = grp / NSB, rs = (grp % NSB) * TRS;
  const int x0 = xp * 2;
  const int hl = g.hd < OLT_CC ? (cg * 4) / g.hd : 0;
  const int b = b0 + img;
  const bf16* tcol = tile + ((img * (TRT + 2) + rs) * (W_ + 2) + x0) * OLT_CC + cg * 4;
  const float* wt = A_s + (((img * TRT + rs) * W_ + x0) * hpc + hl) * 12;
  bf16* out = y + (((long long)b * g.H + r0 + rs) * Wi + w0 + x0) * g.C + c0 + cg * 4;
  const int rows_left = g.H - (r0 + rs);
  const bool bok = b < g.B;
  const long long row_stride = (long long)Wi * g.C;
  const int wps = hpc * 12;
  auto getw = [&](int j, int p, float (&av)[9]) {
    const float* a = wt + (j * W_ + p) * wps;
    const float4 a0 = *reinterpret_cast<const float4*>(a);
    const float4 a1 = *reinterpret_cast<const float4*>(a + 4);
    av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
    av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
    av[8] = a[8];
  };
  olt_walk<W_, TRS>(tcol, getw, [&](int j, const f32x2 (&acc)[2][2]) {
    if (bok && j < rows_left) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        float f0, f1, f2, f3;
        unpk2(acc[p][0], f0, f1);
        unpk2(acc[p][1], f2, f3);
        *reinterpret_cast<uint2*>(out + j * row_stride + p * g.C) = make_uint2(olt_pack_bf16(f0, f1), olt_pack_bf16(f2, f3));
      }
    }
  });
}

template <typename T>
int olt_tmap(CUtensorMap* tm, const void* ptr, long long ld, int B, int H, int W, int tile_w, int rows_box, int NI) {
  const unsigned long long es = sizeof(T);
  unsigned long long dims[4] = {(unsigned long long)ld, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
  unsigned long long str[3] = {(unsigned long long)ld * es, (unsigned long long)W * ld * es, (unsigned long long)H * W * ld * es};
  unsigned box[4] = {(unsigned)OLT_CC, (unsigned)(tile_w + 2), (unsigned)rows_box, (unsigned)NI};
  return ogv_make_tmap(tm, ptr, OGV_BF16, 4, dims, str, box, 0);
}

// the tiled kernels cover bf16, W in {4, 8, 16} or a multiple of 32 (cut into 32-column tiles), 64-channel chunks with
// whole (or part of one) heads, 16-byte rows
inline bool olt_supported(int dtype, int W, int C, int heads, long long ld, const void* a, const void* b) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("OGV_OUTLOOK_TILED"); off = (e && e[0] == '0') ? 1 : 0; }
  if (off || dtype != OGV_BF16) return false;
  if (!(W == 4 || W == 8 || W == 16 || (W >= 32 && W % 32 == 0))) return false;
  const int hd = C / heads;
  if (C % OLT_CC || hd % 4 || !(OLT_CC % hd == 0 || hd % OLT_CC == 0)) return false;
  if (ld % 8 || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) return false;
  return true;
}

template <int W_>
int olt_fwd_launch(const void* va, long long ld, void* y, int B, int H, int W, int C, int heads, cudaStream_t st) {
  using S = OltShape<W_>;
  constexpr int TRT = S::TRS * S::NSB;
  OltGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.hd = C / heads; g.ld = ld;
  g.hpc = g.hd >= OLT_CC ? 1 : OLT_CC / g.hd;
  g.TRT = TRT; g.NI = S::NI; g.NSB = S::NSB;
  g.bands = (H + TRT - 1) / TRT;
  g.nchunks = C / OLT_CC;
  g.tiles_w = W / W_;
  const long long ctas = (long long)((B + S::NI - 1) / S::NI) * g.bands * g.nchunks * g.tiles_w;
  if (ctas > 0x7fffffffLL) { ogv_set_error("outlook_core_fwd: grid too large"); return OGV_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)S::NI * (TRT + 2) * (W_ + 2) * OLT_CC * 2 + (size_t)S::NI * TRT * W_ * g.hpc * 12 * 4;
  CUtensorMap tm;
  if (int rc = olt_tmap<bf16>(&tm, va, ld, B, H, W, W_, TRT + 2, S::NI)) return rc;
  if (int rc = ol_optin(outlook_fwd_tile_kernel<W_>, smem)) return rc;
  outlook_fwd_tile_kernel<W_><<<(unsigned)ctas, OL_THREADS, smem, st>>>(tm, reinterpret_cast<const bf16*>(va),
                                                                       reinterpret_cast<bf16*>(y), g);
  return ogv_check_launch("outlook_core_fwd");
}

// ------------------------------------------------------------------------------------------------
// backward, tiled.  Shared memory: v halo tile | dy halo tile | probabilities of the HALO positions (zero outside the
// image).  Pass A (dv): the forward walker over the dy tile, weight of window element (i, k) = A[q + (i-1, k-1)][8 - (3i+k)]
// (the source position p = q - d_t used tap t to read q).  Pass B: one thread per (position, head) walks the head's
// channels four at a time (start channel staggered across lanes: rows of the tile are 128 bytes apart, the same banks),
// dots dy[p] against the nine v neighbours, finishes the softmax Jacobian and writes the nine logit gradients.
// ------------------------------------------------------------------------------------------------
template <int W_> struct OltShapeB;
template <> struct OltShapeB<32> { static constexpr int NI = 1, NSB = 1, TRS = 4; };
template <> struct OltShapeB<16> { static constexpr int NI = 1, NSB = 2, TRS = 4; };
template <> struct OltShapeB<8> { static constexpr int NI = 2, NSB = 2, TRS = 4; };
template <> struct OltShapeB<4> { static constexpr int NI = 4, NSB = 2, TRS = 2; };

template <int W_>
__global__ void __launch_bounds__(OL_THREADS) outlook_bwd_tile_kernel(const __grid_constant__ CUtensorMap tm_v,
                                                                     const __grid_constant__ CUtensorMap tm_g,
                                                                     const bf16* __restrict__ va, bf16* __restrict__ dva,
                                                                     const OltGeom g) {
  using S = OltShapeB<W_>;
  constexpr int NI = S::NI, NSB = S::NSB, TRS = S::TRS, TRT = TRS * NSB;
  constexpr int W2 = W_ + 2, TR2 = TRT + 2;
  constexpr int TILE_ELEMS = NI * TR2 * W2 * OLT_CC;
  extern __shared__ __align__(128) uint8_t olt_sm[];
  bf16* const vt = reinterpret_cast<bf16*>(olt_sm);
  bf16* const gt = vt + TILE_ELEMS;
  float* const A_s = reinterpret_cast<float*>(olt_sm + 2 * TILE_ELEMS * sizeof(bf16));  // [NI*TR2*W2][hpc][12]
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % g.nchunks;
  int rest = blockIdx.x / g.nchunks;
  const int w0 = (rest % g.tiles_w) * W_;
  rest /= g.tiles_w;
  const int band = rest % g.bands;
  const int b0 = (rest / g.bands) * NI, r0 = band * TRT, c0 = chunk * OLT_CC;
  const int head0 = c0 / g.hd, hpc = g.hpc;
  const int Wi = g.W;
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(&bar, (uint32_t)(2 * TILE_ELEMS * sizeof(bf16)));
    ptx::tma_load_4d(gt, &tm_g, &bar, c0, w0 - 1, r0 - 1, b0);
    ptx::tma_load_4d(vt, &tm_v, &bar, c0, w0 - 1, r0 - 1, b0);
  }
  // probabilities of every halo position of the tile (zero outside the image)
  const int nsm = NI * TR2 * W2 * hpc;
  for (int i = tid; i < nsm; i += OL_THREADS) {
    const int hl = i % hpc, pos = i / hpc;
    const int wc = pos % W2, hr = (pos / W2) % TR2, img = pos / (W2 * TR2);
    const int b = b0 + img, h = r0 - 1 + hr, w = w0 + wc - 1;
    float a[9];
    if (b < g.B && h >= 0 && h < g.H && w >= 0 && w < Wi) {
      const bf16* lp = va + (((long long)b * g.H + h) * Wi + w) * g.ld + g.C + (head0 + hl) * 9;
      float l[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) l[t] = ld1(lp + t);
      softmax9(l, a);
    } else {
#pragma unroll
      for (int t = 0; t < 9; ++t) a[t] = 0.f;
    }
    float* dst = A_s + i * 12;
    *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(a[4], a[5], a[6], a[7]);
    dst[8] = a[8];
  }
  __syncthreads();
  ptx::mbar_wait(&bar, 0);

  // ---- pass A: dv ----
  {
    const int cg = tid % 16;
    const int xp = (tid / 16) % (W_ / 2);
    const int grp = tid / (8 * W_);
    const int img = grp / NSB, rs = (grp % NSB) * TRS;
    const int x0 = xp * 2;
    const int hl = g.hd < OLT_CC ? (cg * 4) / g.hd : 0;
    const int b = b0 + img;
    const bf16* tcol = gt + ((img * TR2 + rs) * W2 + x0) * OLT_CC + cg * 4;
    // halo position (row rs + j + i, column x0 + p + k) <-> q + (i - 1, k - 1)
    const float* abase = A_s + (((img * TR2 + rs) * W2 + x0) * hpc + hl) * 12;
    const int aps = hpc * 12;
    auto getw = [&](int j, int p, float (&av)[9]) {
#pragma unroll
      for (int s9 = 0; s9 < 9; ++s9) av[s9] = abase[((j + s9 / 3) * W2 + p + s9 % 3) * aps + (8 - s9)];
    };
    bf16* out = dva + (((long long)b * g.H + r0 + rs) * Wi + w0 + x0) * g.ld + c0 + cg * 4;
    const int rows_left = g.H - (r0 + rs);
    const bool bok = b < g.B;
    const long long row_stride = (long long)Wi * g.ld;
    olt_walk<W_, TRS>(tcol, getw, [&](int j, const f32x2 (&acc)[2][2]) {
      if (bok && j < rows_left) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float f0, f1, f2, f3;
          unpk2(acc[p][0], f0, f1);
          unpk2(acc[p][1], f2, f3);
          *reinterpret_cast<uint2*>(out + j * row_stride + p * g.ld) =
              make_uint2(olt_pack_bf16(f0, f1), olt_pack_bf16(f2, f3));
        }
      }
    });
  }

  // ---- pass B: dA[p, head, t] = sum_c dy[p, c] * v[p + d_t, c];  dlogits = A * (dA - sum_t A * dA) ----
  const int nitems = NI * TRT * W_ * hpc;
  const int hdc = g.hd < OLT_CC ? g.hd : OLT_CC;  // channels of the head inside this chunk
  const int ncg = hdc / 4;
  const int nl = g.heads * 9;
  for (int i = tid; i < nitems; i += OL_THREADS) {
    const int hl = i % hpc, pos = i / hpc;
    const int w = pos % W_, rr = (pos / W_) % TRT, img = pos / (W_ * TRT);
    const int b = b0 + img, h = r0 + rr;
    if (b >= g.B || h >= g.H) continue;
    const int centre = ((img * TR2 + rr + 1) * W2 + w + 1);
    const bf16* vq = vt + centre * OLT_CC + hl * hdc;
    const bf16* gq = gt + centre * OLT_CC + hl * hdc;
    f32x2 d2[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) d2[t] = 0ull;
    const int rot = pos & (ncg - 1);  // ncg is a power of two (hd in {4, 8, 16, 32, 64})
    for (int it = 0; it < ncg; ++it) {
      const int co = ((it + rot) & (ncg - 1)) * 4;
      const uint2 gu = *reinterpret_cast<const uint2*>(gq + co);
      const f32x2 g0 = pk2(__uint_as_float(gu.x << 16), __uint_as_float(gu.x & 0xffff0000u));
      const f32x2 g1 = pk2(__uint_as_float(gu.y << 16), __uint_as_float(gu.y & 0xffff0000u));
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint2 vu = *reinterpret_cast<const uint2*>(vq + ((t / 3 - 1) * W2 + (t % 3 - 1)) * OLT_CC + co);
        d2[t] = fma2(g0, pk2(__uint_as_float(vu.x << 16), __uint_as_float(vu.x & 0xffff0000u)), d2[t]);
        d2[t] = fma2(g1, pk2(__uint_as_float(vu.y << 16), __uint_as_float(vu.y & 0xffff0000u)), d2[t]);
      }
    }
    const float* a = A_s + (centre * hpc + hl) * 12;
    float dA[9], dot = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float lo, hi;
      unpk2(d2[t], lo, hi);
      dA[t] = lo + hi;
      dot = fmaf(a[t], dA[t], dot);
    }
    bf16* row = dva + (((long long)b * g.H + h) * Wi + w0 + w) * g.ld;
    bf16* out = row + g.C + (head0 + hl) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) st1(out + t, a[t] * (dA[t] - dot));
    if (chunk == 0 && hl == 0) ol_zero_pad(row, g.C + nl, (int)g.ld);
  }
}

template <int W_>
int olt_bwd_launch(const void* va, long long ld, const void* dy, void* dva, int B, int H, int W, int C, int heads,
                   cudaStream_t st) {
  using S = OltShapeB<W_>;
  constexpr int TRT = S::TRS * S::NSB;
  OltGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.hd = C / heads; g.ld = ld;
  g.hpc = g.hd >= OLT_CC ? 1 : OLT_CC / g.hd;
  g.TRT = TRT; g.NI = S::NI; g.NSB = S::NSB;
  g.bands = (H + TRT - 1) / TRT;
  g.nchunks = C / OLT_CC;
  g.tiles_w = W / W_;
  const long long ctas = (long long)((B + S::NI - 1) / S::NI) * g.bands * g.nchunks * g.tiles_w;
  if (ctas > 0x7fffffffLL) { ogv_set_error("outlook_core_bwd: grid too large"); return OGV_ERR_UNSUPPORTED; }
  const size_t tile = (size_t)S::NI * (TRT + 2) * (W_ + 2) * OLT_CC * 2;
  const size_t smem = 2 * tile + (size_t)S::NI * (TRT + 2) * (W_ + 2) * g.hpc * 12 * 4;
  CUtensorMap tmv, tmg;
  if (int rc = olt_tmap<bf16>(&tmv, va, ld, B, H, W, W_, TRT + 2, S::NI)) return rc;
  if (int rc = olt_tmap<bf16>(&tmg, dy, C, B, H, W, W_, TRT + 2, S::NI)) return rc;
  if (int rc = ol_optin(outlook_bwd_tile_kernel<W_>, smem)) return rc;
  outlook_bwd_tile_kernel<W_><<<(unsigned)ctas, OL_THREADS, smem, st>>>(tmv, tmg, reinterpret_cast<const bf16*>(va),
                                                                       reinterpret_cast<bf16*>(dva), g);
  return ogv_check_launch("outlook_core_bwd");
}

}  // namespace

#define OL_DISPATCH_VEC(vec, ...)                          \
  do {                                                     \
    if ((vec) == 8) { constexpr int VEC = 8; __VA_ARGS__; } \
    else if ((vec) == 4) { constexpr int VEC = 4; __VA_ARGS__; } \
    else { constexpr int VEC = 1; __VA_ARGS__; }           \
  } while (0)

extern "C" int ogv_outlook_core_fwd(const void* va, long long ld_va, void* y, int B, int H, int W, int C, int heads,
                                    int dtype, void* stream) {
  if ((long long)B * H * W == 0) return OGV_OK;
  OGV_REQUIRE(va && y, "outlook_core_fwd: null pointer");
  OGV_REQUIRE(heads > 0 && C > 0 && C % heads == 0, "outlook_core_fwd: dim must be divisible by num_heads");
  OGV_REQUIRE(ld_va >= C + 9 * heads, "outlook_core_fwd: ld_va=%lld < C+9*heads", ld_va);
  cudaStream_t st = (cudaStream_t)stream;
  if (olt_supported(dtype, W, C, heads, ld_va, va, y)) {
    switch (W) {
      case 16: return olt_fwd_launch<16>(va, ld_va, y, B, H, W, C, heads, st);
      case 8: return olt_fwd_launch<8>(va, ld_va, y, B, H, W, C, heads, st);
      case 4: return olt_fwd_launch<4>(va, ld_va, y, B, H, W, C, heads, st);
      default: return olt_fwd_launch<32>(va, ld_va, y, B, H, W, C, heads, st);
    }
  }
  OlGeom g;
  const int vec = pick_vec(C, C / heads, ld_va);
  if (ol_geom(B, H, W, C, heads, ld_va, vec, &g)) { ogv_set_error("outlook_core_fwd: image %dx%d too wide", H, W); return OGV_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)g.TR * W * heads * 9 * sizeof(float);
  OGV_DISPATCH_DTYPE(dtype, T, {
    OL_DISPATCH_VEC(vec, {
      if (int rc = ol_optin(outlook_fwd_kernel<T, VEC>, smem)) return rc;
      outlook_fwd_kernel<T, VEC><<<B * g.bands, OL_THREADS, smem, st>>>(reinterpret_cast<const T*>(va),
                                                                        reinterpret_cast<T*>(y), g);
    });
    return ogv_check_launch("outlook_core_fwd");
  });
}

extern "C" int ogv_outlook_core_bwd(const void* va, long long ld_va, const void* dy, void* dva, int B, int H, int W,
                                    int C, int heads, int dtype, void* stream) {
  if ((long long)B * H * W == 0) return OGV_OK;
  OGV_REQUIRE(va && dy && dva, "outlook_core_bwd: null pointer");
  OGV_REQUIRE(heads > 0 && C > 0 && C % heads == 0, "outlook_core_bwd: dim must be divisible by num_heads");
  OGV_REQUIRE(ld_va >= C + 9 * heads, "outlook_core_bwd: ld_va=%lld < C+9*heads", ld_va);
  cudaStream_t st = (cudaStream_t)stream;
  // the tiled backward finishes a head's logit gradients inside one CTA: heads no wider than the 64-channel chunk
  if (olt_supported(dtype, W, C, heads, ld_va, va, dva) && C / heads <= OLT_CC && C % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    switch (W) {
      case 16: return olt_bwd_launch<16>(va, ld_va, dy, dva, B, H, W, C, heads, st);
      case 8: return olt_bwd_launch<8>(va, ld_va, dy, dva, B, H, W, C, heads, st);
      case 4: return olt_bwd_launch<4>(va, ld_va, dy, dva, B, H, W, C, heads, st);
      default: return olt_bwd_launch<32>(va, ld_va, dy, dva, B, H, W, C, heads, st);
    }
  }
  OlGeom g;
  const int vec = pick_vec(C, C / heads, ld_va);
  if (ol_geom(B, H, W, C, heads, ld_va, vec, &g)) { ogv_set_error("outlook_core_bwd: image %dx%d too wide", H, W); return OGV_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)(2 * g.TR + 2) * W * heads * 9 * sizeof(float);
  OGV_DISPATCH_DTYPE(dtype, T, {
    // shuffle reduction needs the hd/VEC threads of a head to be an aligned power-of-two lane group
    const int gs = g.hd / vec;
    const bool shfl = gs >= 1 && gs <= 32 && (gs & (gs - 1)) == 0;
    OL_DISPATCH_VEC(vec, {
      if (shfl) {
        if (int rc = ol_optin(outlook_bwd_kernel<T, VEC, true>, smem)) return rc;
        outlook_bwd_kernel<T, VEC, true><<<B * g.bands, OL_THREADS, smem, st>>>(
            reinterpret_cast<const T*>(va), reinterpret_cast<const T*>(dy), reinterpret_cast<T*>(dva), g);
      } else {
        if (int rc = ol_optin(outlook_bwd_kernel<T, VEC, false>, smem)) return rc;
        outlook_bwd_kernel<T, VEC, false><<<B * g.bands, OL_THREADS, smem, st>>>(
            reinterpret_cast<const T*>(va), reinterpret_cast<const T*>(dy), reinterpret_cast<T*>(dva), g);
      }
    });
    return ogv_check_launch("outlook_core_bwd");
  });
}
