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
#include "ogv_common.cuh"
#include "ogv_tma.cuh"  // FastDiv
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
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      ogv_set_error("outlook: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return OGV_ERR_CUDA;
    }
  }
  return OGV_OK;
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
  OlGeom g;
  const int vec = pick_vec(C, C / heads, ld_va);
  if (ol_geom(B, H, W, C, heads, ld_va, vec, &g)) { ogv_set_error("outlook_core_fwd: image %dx%d too wide", H, W); return OGV_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)g.TR * W * heads * 9 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
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
  OlGeom g;
  const int vec = pick_vec(C, C / heads, ld_va);
  if (ol_geom(B, H, W, C, heads, ld_va, vec, &g)) { ogv_set_error("outlook_core_bwd: image %dx%d too wide", H, W); return OGV_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)(2 * g.TR + 2) * W * heads * 9 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
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
