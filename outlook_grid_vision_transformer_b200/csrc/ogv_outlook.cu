// Outlook attention core: softmax over the k*k=9 logits of each head, then the weighted gather over
// the 3x3 neighbourhood of v (zero outside the image, weights NOT renormalised) -- the reference's
// F.unfold + mul + sum (outlook_attention.py:104-120) without ever materialising the 9x im2col tensor.
// Backward is the overlap-add (col2im) written as a gather plus the softmax Jacobian.
//
// `va` rows: [ v (C) | logits head0 (9) | logits head1 (9) | ... | zero padding ] with stride ld_va.
#include "ogv_common.cuh"
#include "../../include/ogv.h"

namespace {

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
  float inv = 1.f / s;
#pragma unroll
  for (int t = 0; t < 9; ++t) a[t] *= inv;
}

template <typename T>
__device__ __forceinline__ void load_logits(const T* row, int C, int head, float (&l)[9]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) l[t] = ld1(row + C + head * 9 + t);
}

// one thread per (position, VEC-channel group)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) outlook_fwd_kernel(const T* __restrict__ va, long long ld, T* __restrict__ y,
                                                          int B, int H, int W, int C, int hd) {
  const int nvc = C / VEC;
  const long long total = (long long)B * H * W * nvc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % nvc);
    const long long m = idx / nvc;
    const int w = (int)(m % W);
    const int h = (int)((m / W) % H);
    const int c = cg * VEC;
    const int head = c / hd;
    float l[9], a[9];
    load_logits(va + m * ld, C, head, l);
    softmax9(l, a);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        float v[VEC];
        ldv<VEC>(va + (m + (long long)(t / 3 - 1) * W + (t % 3 - 1)) * ld + c, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(a[t], v[i], acc[i]);
      }
    }
    stv<VEC>(y + m * C + c, acc);
  }
}

// dv[q,c] = sum_t A[head, t, q - d_t] * dy[q - d_t, c]
template <typename T, int VEC>
__global__ void __launch_bounds__(256) outlook_bwd_dv_kernel(const T* __restrict__ va, long long ld,
                                                             const T* __restrict__ dy, T* __restrict__ dva, int B,
                                                             int H, int W, int C, int hd) {
  const int nvc = C / VEC;
  const long long total = (long long)B * H * W * nvc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % nvc);
    const long long m = idx / nvc;
    const int w = (int)(m % W);
    const int h = (int)((m / W) % H);
    const int c = cg * VEC;
    const int head = c / hd;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      // source position p = q - d_t ; it used tap t to read q
      const int hh = h - (t / 3 - 1), ww = w - (t % 3 - 1);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const long long p = m - (long long)(t / 3 - 1) * W - (t % 3 - 1);
        float l[9], a[9];
        load_logits(va + p * ld, C, head, l);
        softmax9(l, a);
        float g[VEC];
        ldv<VEC>(dy + p * C + c, g);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(a[t], g[i], acc[i]);
      }
    }
    stv<VEC>(dva + m * ld + c, acc);
  }
}

// one thread per (position, head): dA[t] = sum_d dy[p,d] v[p+d_t,d]; dlogits = A*(dA - sum A*dA)
template <typename T, int VEC>
__global__ void __launch_bounds__(128) outlook_bwd_dl_kernel(const T* __restrict__ va, long long ld,
                                                             const T* __restrict__ dy, T* __restrict__ dva, int B,
                                                             int H, int W, int C, int heads, int hd) {
  const long long total = (long long)B * H * W * heads;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int head = (int)(idx % heads);
    const long long m = idx / heads;
    const int w = (int)(m % W);
    const int h = (int)((m / W) % H);
    float l[9], a[9], dA[9];
    load_logits(va + m * ld, C, head, l);
    softmax9(l, a);
#pragma unroll
    for (int t = 0; t < 9; ++t) dA[t] = 0.f;
    const int c0 = head * hd;
    for (int d = 0; d < hd; d += VEC) {
      float g[VEC];
      ldv<VEC>(dy + m * C + c0 + d, g);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          float v[VEC];
          ldv<VEC>(va + (m + (long long)(t / 3 - 1) * W + (t % 3 - 1)) * ld + c0 + d, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) dA[t] = fmaf(g[i], v[i], dA[t]);
        }
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) dot = fmaf(a[t], dA[t], dot);
    T* out = dva + m * ld + C + head * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) st1(out + t, a[t] * (dA[t] - dot));
    if (head == 0) {  // zero the alignment padding so the dgrad/wgrad GEMMs see exact zeros
      for (int j = C + heads * 9; j < ld; ++j) st1(dva + m * ld + j, 0.f);
    }
  }
}

inline int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)ogv_num_sms() * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

inline int pick_vec(int C, int hd, long long ld) {
  if (hd % 8 == 0 && C % 8 == 0 && ld % 8 == 0) return 8;
  if (hd % 4 == 0 && C % 4 == 0 && ld % 4 == 0) return 4;
  return 1;
}

}  // namespace

extern "C" int ogv_outlook_core_fwd(const void* va, long long ld_va, void* y, int B, int H, int W, int C, int heads,
                                    int dtype, void* stream) {
  OGV_REQUIRE(va && y, "outlook_core_fwd: null pointer");
  OGV_REQUIRE(heads > 0 && C > 0 && C % heads == 0, "outlook_core_fwd: dim must be divisible by num_heads");
  OGV_REQUIRE(ld_va >= C + 9 * heads, "outlook_core_fwd: ld_va=%lld < C+9*heads", ld_va);
  const int hd = C / heads;
  const int vec = pick_vec(C, hd, ld_va);
  const long long total = (long long)B * H * W * (C / vec);
  if (total == 0) return OGV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(total, 256);
  OGV_DISPATCH_DTYPE(dtype, T, {
    const T* p = reinterpret_cast<const T*>(va);
    T* o = reinterpret_cast<T*>(y);
    if (vec == 8) outlook_fwd_kernel<T, 8><<<grid, 256, 0, st>>>(p, ld_va, o, B, H, W, C, hd);
    else if (vec == 4) outlook_fwd_kernel<T, 4><<<grid, 256, 0, st>>>(p, ld_va, o, B, H, W, C, hd);
    else outlook_fwd_kernel<T, 1><<<grid, 256, 0, st>>>(p, ld_va, o, B, H, W, C, hd);
    return ogv_check_launch("outlook_core_fwd");
  });
}

extern "C" int ogv_outlook_core_bwd(const void* va, long long ld_va, const void* dy, void* dva, int B, int H, int W,
                                    int C, int heads, int dtype, void* stream) {
  OGV_REQUIRE(va && dy && dva, "outlook_core_bwd: null pointer");
  OGV_REQUIRE(heads > 0 && C > 0 && C % heads == 0, "outlook_core_bwd: dim must be divisible by num_heads");
  OGV_REQUIRE(ld_va >= C + 9 * heads, "outlook_core_bwd: ld_va=%lld < C+9*heads", ld_va);
  const int hd = C / heads;
  const int vec = pick_vec(C, hd, ld_va);
  const long long total = (long long)B * H * W * (C / vec);
  const long long total_dl = (long long)B * H * W * heads;
  if (total == 0) return OGV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, {
    const T* p = reinterpret_cast<const T*>(va);
    const T* g = reinterpret_cast<const T*>(dy);
    T* o = reinterpret_cast<T*>(dva);
    const int g1 = grid_for(total, 256), g2 = grid_for(total_dl, 128);
    if (vec == 8) {
      outlook_bwd_dv_kernel<T, 8><<<g1, 256, 0, st>>>(p, ld_va, g, o, B, H, W, C, hd);
      outlook_bwd_dl_kernel<T, 8><<<g2, 128, 0, st>>>(p, ld_va, g, o, B, H, W, C, heads, hd);
    } else if (vec == 4) {
      outlook_bwd_dv_kernel<T, 4><<<g1, 256, 0, st>>>(p, ld_va, g, o, B, H, W, C, hd);
      outlook_bwd_dl_kernel<T, 4><<<g2, 128, 0, st>>>(p, ld_va, g, o, B, H, W, C, heads, hd);
    } else {
      outlook_bwd_dv_kernel<T, 1><<<g1, 256, 0, st>>>(p, ld_va, g, o, B, H, W, C, hd);
      outlook_bwd_dl_kernel<T, 1><<<g2, 128, 0, st>>>(p, ld_va, g, o, B, H, W, C, heads, hd);
    }
    return ogv_check_launch("outlook_core_bwd");
  });
}
