// LayerNorm (per position over channels) and BatchNorm2d (per channel over B*H*W) kernels.
// Reference: outlook_attention.py:17-31 (LayerNorm2d), Out_Grid_Block.py:69,84 (nn.LayerNorm),
// mbc_conv.py:61 (nn.BatchNorm2d, batch statistics in training).
#include "ogv_common.cuh"
#include "ogv_reduce.cuh"
#include "../../include/ogv.h"

// row-groups in flight per warp iteration (U) and resident CTAs per SM the kernels are compiled for
#ifndef LN_FWD_U
#define LN_FWD_U 2
#endif
#ifndef LN_BWD_U
#define LN_BWD_U 2
#endif
#ifndef LN_FWD_MINB
#define LN_FWD_MINB 4
#endif
#ifndef LN_BWD_MINB
#define LN_BWD_MINB 2
#endif

namespace {

// ------------------------------------------------------------------ LayerNorm
// A row (C channels = nv 16-byte vectors) is owned by a group of G lanes, VPL vectors per lane, so
// that all 32 lanes of a warp move data even at C = 64 (G = 8 -> 4 rows per warp).
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// U row-groups per warp iteration: the loads of all U rows are issued first (kept in storage format), then the
// rows are normalised one after the other -- bytes in flight per SM, not arithmetic, bound these kernels.
template <typename T, int G, int VPL, int U>
__global__ void __launch_bounds__(256, VPL == 1 ? LN_FWD_MINB : 1) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     long long M, int C, float eps) {
  constexpr int RPW = 32 / G;  // rows per warp and row-group
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gr = lane / G;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  const int nv = C >> 3;
  const float inv_c = 1.f / (float)C;
  float g[VPL][8], b[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int idx = gl + G * j;
    if (idx < nv) {
      ld8(gamma + idx * 8, g[j]);
      ld8(beta + idx * 8, b[j]);
    }
  }
  for (long long base = warp0 * (RPW * U); base < M; base += nwarps * (RPW * U)) {
    Raw8<T> rx[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = base + u * RPW + gr;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        raw8_zero(rx[u][j]);
        if (row < M && idx < nv) ld_raw8(x + row * C + idx * 8, rx[u][j]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = base + u * RPW + gr;
      const bool rv = row < M;
      float v[VPL][8];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        cvt_raw8(rx[u][j], v[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[j][i];
      }
      const float mean = group_sum<G>(s) * inv_c;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        if (idx < nv) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float d = v[j][i] - mean;
            q += d * d;
          }
        }
      }
      const float rstd = 1.f / sqrtf(group_sum<G>(q) * inv_c + eps);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        if (rv && idx < nv) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = (v[j][i] - mean) * rstd * g[j][i] + b[j][i];
          st8(y + row * C + idx * 8, o);
        }
      }
      if (rv && gl == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
      }
    }
  }
}

template <typename T, int G, int VPL, int U>
__global__ void __launch_bounds__(256, VPL == 1 ? LN_BWD_MINB : 1)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ dres,
              T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C) {
  constexpr int RPW = 32 / G;
  __shared__ float sg[1024];
  __shared__ float sb[1024];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sg[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gr = lane / G;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  const int nv = C >> 3;
  const float inv_c = 1.f / (float)C;
  float ag[VPL][8], ab[VPL][8], gm[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int idx = gl + G * j;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[j][i] = 0.f; ab[j][i] = 0.f; gm[j][i] = 0.f; }
    if (idx < nv) ld8(gamma + idx * 8, gm[j]);
  }
  for (long long base = warp0 * (RPW * U); base < M; base += nwarps * (RPW * U)) {
    // every load of the U rows (gradient, input, residual gradient, row statistics) is in flight before any use
    Raw8<T> rd[U][VPL], rx[U][VPL], rr[U][VPL];
    float mu[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = base + u * RPW + gr;
      const bool rv = row < M;
      mu[u] = rv ? mean[row] : 0.f;
      rs[u] = rv ? rstd[row] : 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        raw8_zero(rd[u][j]);
        raw8_zero(rx[u][j]);
        raw8_zero(rr[u][j]);
        if (rv && idx < nv) {
          ld_raw8(dy + row * C + idx * 8, rd[u][j]);
          ld_raw8(x + row * C + idx * 8, rx[u][j]);
          if (dres) ld_raw8(dres + row * C + idx * 8, rr[u][j]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = base + u * RPW + gr;
      const bool rv = row < M;
      float gg[VPL][8], xh[VPL][8];
      float c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        float d[8], xv[8];
        cvt_raw8(rd[u][j], d);  // zeros outside the row / beyond C: no contribution below
        cvt_raw8(rx[u][j], xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[j][i] = (xv[i] - mu[u]) * rs[u];
          ag[j][i] += d[i] * xh[j][i];
          ab[j][i] += d[i];
          gg[j][i] = d[i] * gm[j][i];
          c1 += gg[j][i];
          c2 += gg[j][i] * xh[j][i];
        }
      }
      c1 = group_sum<G>(c1) * inv_c;
      c2 = group_sum<G>(c2) * inv_c;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        if (rv && idx < nv) {
          float o[8], r[8];
          cvt_raw8(rr[u][j], r);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = rs[u] * (gg[j][i] - c1 - xh[j][i] * c2) + r[i];
          st8(dx + row * C + idx * 8, o);
        }
      }
    }
  }
  // fold the RPW row-groups of the warp, then one shared-memory atomic per channel per warp
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int idx = gl + G * j;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = ag[j][i], b = ab[j][i];
#pragma unroll
      for (int o = 16; o >= G; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (gr == 0 && idx < nv) {
        atomicAdd(&sg[idx * 8 + i], a);
        atomicAdd(&sb[idx * 8 + i], b);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sg[i]);
    if (dbeta) atomicAdd(dbeta + i, sb[i]);
  }
}

#define LN_DISPATCH(C, ...)                                  \
  do {                                                        \
    if ((C) <= 32) { constexpr int G = 4, VPL = 1; __VA_ARGS__; }    \
    else if ((C) <= 64) { constexpr int G = 8, VPL = 1; __VA_ARGS__; }  \
    else if ((C) <= 128) { constexpr int G = 16, VPL = 1; __VA_ARGS__; } \
    else if ((C) <= 256) { constexpr int G = 32, VPL = 1; __VA_ARGS__; } \
    else if ((C) <= 512) { constexpr int G = 32, VPL = 2; __VA_ARGS__; } \
    else { constexpr int G = 32, VPL = 4; __VA_ARGS__; }             \
  } while (0)

// ------------------------------------------------------------------ BatchNorm
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float n, int C, float eps, float momentum,
                                   int training) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    mean = sum[c] / n;
    var = fmaxf(sumsq[c] / n - mean * mean, 0.f);
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) {
      float unbiased = n > 1.f ? var * n / (n - 1.f) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  float rstd = 1.f / sqrtf(var + eps);
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * rstd;
  shift[c] = b - mean * g * rstd;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                const float* __restrict__ shift, const T* __restrict__ res, T* __restrict__ out,
                                long long nvec, int nv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % nv);
    float v[8], sc[8], sh[8];
    ld8(x + i * 8, v);
    ld8(scale + cv * 8, sc);
    ld8(shift + cv * 8, sh);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = v[k] * sc[k] + sh[k];
    if (res) {
      float r[8];
      ld8(res + i * 8, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += r[k];
    }
    st8(out + i * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(COLREDUCE_THREADS, 2) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int nv) {
  float acc[2][8];
  colreduce_init(acc);
  float mu[8], rs[8];
  {
    const int cv0 = threadIdx.x % nv;
    ld8(mean + cv0 * 8, mu);
    ld8(rstd + cv0 * 8, rs);
  }
  COLREDUCE_LOOP(M, nv, row, cv) {
    float d[8], xv[8];
    ld8(dy + (row * nv + cv) * 8, d);
    ld8(x + (row * nv + cv) * 8, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += d[i];
      acc[1][i] += d[i] * (xv[i] - mu[i]) * rs[i];
    }
  }
  float* outs[2] = {dbeta, dgamma};
  colreduce_finish<2>(acc, outs, nv);
}

// dx = gamma*rstd*(dy - dbeta/n - xhat*dgamma/n).  A thread owns one channel vector (cv) and strides over rows, so
// the five per-channel parameter vectors are folded ONCE into  dx = A*dy - (Cx*x + Bc)  (they used to be re-read,
// with a 64-bit modulo, for every 16 bytes of data); U rows are in flight per thread in storage format.
template <typename T, int U>
__global__ void __launch_bounds__(COLREDUCE_THREADS)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dgamma,
                    const float* __restrict__ dbeta, T* __restrict__ dx, long long M, int nv, float inv_n) {
  const int cv = threadIdx.x % nv;
  const int kk = blockDim.x / nv;
  float A[8], Bc[8], Cx[8];
  {
    float mu[8], rs[8], g[8], dg[8], db[8];
    ld8(mean + cv * 8, mu);
    ld8(rstd + cv * 8, rs);
    ld8(gamma + cv * 8, g);
    ld8(dgamma + cv * 8, dg);
    ld8(dbeta + cv * 8, db);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      A[k] = g[k] * rs[k];
      Cx[k] = A[k] * rs[k] * dg[k] * inv_n;
      Bc[k] = A[k] * db[k] * inv_n - Cx[k] * mu[k];
    }
  }
  const long long stride = (long long)gridDim.x * kk;
  for (long long row0 = (long long)blockIdx.x * kk + threadIdx.x / nv; row0 < M; row0 += stride * U) {
    Raw8<T> rd[U], rx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < M) {
        ld_raw8(dy + (row * nv + cv) * 8, rd[u]);
        ld_raw8(x + (row * nv + cv) * 8, rx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < M) {
        float d[8], xv[8], o[8];
        cvt_raw8(rd[u], d);
        cvt_raw8(rx[u], xv);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], d[k], -fmaf(Cx[k], xv[k], Bc[k]));
        st8(dx + (row * nv + cv) * 8, o);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// BatchNorm + activation of the conv → BN → act units around the blocks (stem_head.py:23-32, downsampling.py:28-65):
// out = act(scale*x + shift), and its backward with g = dy * act'(scale*x + shift) recomputed on the fly in both
// passes (nothing but the pre-BN tensor is saved).  Same thread layout as bn_apply / bn_bwd_reduce / bn_bwd_apply.
// ------------------------------------------------------------------------------------------------
template <typename T, int ACT, int U>
__global__ void __launch_bounds__(COLREDUCE_THREADS)
bn_act_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                    T* __restrict__ out, long long M, int nv) {
  constexpr bool FAST = FastAct<T>::value;
  const int cv = threadIdx.x % nv;
  const int kk = blockDim.x / nv;
  float sc[8], sh[8];
  ld8(scale + cv * 8, sc);
  ld8(shift + cv * 8, sh);
  const long long stride = (long long)gridDim.x * kk;
  for (long long row0 = (long long)blockIdx.x * kk + threadIdx.x / nv; row0 < M; row0 += stride * U) {
    Raw8<T> rx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < M) ld_raw8(x + (row * nv + cv) * 8, rx[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < M) {
        float v[8];
        cvt_raw8(rx[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = act_apply_t<ACT, FAST>(fmaf(v[k], sc[k], sh[k]));
        st8(out + (row * nv + cv) * 8, v);
      }
    }
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(COLREDUCE_THREADS, 2)
bn_act_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         long long M, int nv) {
  constexpr bool FAST = FastAct<T>::value;
  float acc[2][8];
  colreduce_init(acc);
  float mu[8], rs[8], sc[8], sh[8];
  {
    const int cv0 = threadIdx.x % nv;
    ld8(mean + cv0 * 8, mu);
    ld8(rstd + cv0 * 8, rs);
    ld8(scale + cv0 * 8, sc);
    ld8(shift + cv0 * 8, sh);
  }
  COLREDUCE_LOOP(M, nv, row, cv) {
    float d[8], xv[8];
    ld8(dy + (row * nv + cv) * 8, d);
    ld8(x + (row * nv + cv) * 8, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = d[i] * act_grad_t<ACT, FAST>(fmaf(xv[i], sc[i], sh[i]));
      acc[0][i] += g;
      acc[1][i] += g * (xv[i] - mu[i]) * rs[i];
    }
  }
  float* outs[2] = {dbeta, dgamma};
  colreduce_finish<2>(acc, outs, nv);
}

template <typename T, int ACT, int U>
__global__ void __launch_bounds__(COLREDUCE_THREADS)
bn_act_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ scale,
                        const float* __restrict__ shift, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                        const float* __restrict__ dgamma, const float* __restrict__ dbeta, T* __restrict__ dx,
                        long long M, int nv, float inv_n) {
  constexpr bool FAST = FastAct<T>::value;
  const int cv = threadIdx.x % nv;
  const int kk = blockDim.x / nv;
  float A[8], Bc[8], Cx[8], sc[8], sh[8];
  {
    float mu[8], rs[8], g[8], dg[8], db[8];
    ld8(mean + cv * 8, mu);
    ld8(rstd + cv * 8, rs);
    ld8(gamma + cv * 8, g);
    ld8(dgamma + cv * 8, dg);
    ld8(dbeta + cv * 8, db);
    ld8(scale + cv * 8, sc);
    ld8(shift + cv * 8, sh);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      A[k] = g[k] * rs[k];
      Cx[k] = A[k] * rs[k] * dg[k] * inv_n;
      Bc[k] = A[k] * db[k] * inv_n - Cx[k] * mu[k];
    }
  }
  const long long stride = (long long)gridDim.x * kk;
  for (long long row0 = (long long)blockIdx.x * kk + threadIdx.x / nv; row0 < M; row0 += stride * U) {
    Raw8<T> rd[U], rx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < M) {
        ld_raw8(dy + (row * nv + cv) * 8, rd[u]);
        ld_raw8(x + (row * nv + cv) * 8, rx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = row0 + u * stride;
      if (row < M) {
        float d[8], xv[8], o[8];
        cvt_raw8(rd[u], d);
        cvt_raw8(rx[u], xv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float g = d[k] * act_grad_t<ACT, FAST>(fmaf(xv[k], sc[k], sh[k]));
          o[k] = fmaf(A[k], g, -fmaf(Cx[k], xv[k], Bc[k]));
        }
        st8(dx + (row * nv + cv) * 8, o);
      }
    }
  }
}

inline int flat_grid(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)ogv_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int ogv_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                 float* rstd, long long M, int C, float eps, int dtype, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null pointer");
  OGV_REQUIRE(C > 0 && C % 8 == 0 && C <= 1024, "layernorm_fwd: C=%d must be a multiple of 8 and <= 1024", C);
  const int ln_g = C <= 32 ? 4 : (C <= 64 ? 8 : (C <= 128 ? 16 : 32));
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, {
    const T* xp = reinterpret_cast<const T*>(x);
    T* yp = reinterpret_cast<T*>(y);
    LN_DISPATCH(C, {
      constexpr int U = LN_FWD_U / VPL > 0 ? LN_FWD_U / VPL : 1;
      const int grid = flat_grid((M * ln_g + U - 1) / U, 256);
      ln_fwd_kernel<T, G, VPL, U><<<grid, 256, 0, st>>>(xp, gamma, beta, yp, mean, rstd, M, C, eps);
    });
    return ogv_check_launch("layernorm_fwd");
  });
}

extern "C" int ogv_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                 const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 long long M, int C, int dtype, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(dy && x && gamma && mean && rstd && dx, "layernorm_bwd: null pointer");
  OGV_REQUIRE(C > 0 && C % 8 == 0 && C <= 1024, "layernorm_bwd: C=%d must be a multiple of 8 and <= 1024", C);
  const int ln_g = C <= 32 ? 4 : (C <= 64 ? 8 : (C <= 128 ? 16 : 32));
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, {
    const T* dyp = reinterpret_cast<const T*>(dy);
    const T* xp = reinterpret_cast<const T*>(x);
    const T* rp = reinterpret_cast<const T*>(dres);
    T* dxp = reinterpret_cast<T*>(dx);
    LN_DISPATCH(C, {
      constexpr int U = (VPL == 1 && sizeof(T) == 2) ? LN_BWD_U : 1;
      // one resident wave of grid-striding CTAs: one dgamma / dbeta atomic flush per CTA
      const long long want = (M * ln_g + 256 * U - 1) / (256 * U);
      const long long cap = (long long)ogv_num_sms() * (VPL == 1 ? LN_BWD_MINB : 1);
      const int grid = (int)(want < cap ? want : cap);
      ln_bwd_kernel<T, G, VPL, U><<<grid, 256, 0, st>>>(dyp, xp, gamma, mean, rstd, rp, dxp, dgamma, dbeta, M, C);
    });
    return ogv_check_launch("layernorm_bwd");
  });
}

extern "C" int ogv_bn_finalize(const float* sum, const float* sumsq, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                               float* rstd, long long n, int C, float eps, float momentum, int training,
                               void* stream) {
  OGV_REQUIRE(scale && shift && C > 0, "bn_finalize: bad args");
  OGV_REQUIRE(training ? (sum && sumsq && n > 0) : (running_mean && running_var),
              "bn_finalize: missing statistics for mode training=%d", training);
  bn_finalize_kernel<<<ogv_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(
      sum, sumsq, gamma, beta, running_mean, running_var, scale, shift, mean, rstd, (float)n, C, eps, momentum,
      training);
  return ogv_check_launch("bn_finalize");
}

extern "C" int ogv_bn_apply(const void* x, const float* scale, const float* shift, const void* res, void* out,
                            long long M, int C, int dtype, void* stream) {
  OGV_REQUIRE(x && scale && shift && out && C % 8 == 0, "bn_apply: bad args (C %% 8 == 0)");
  long long nvec = M * (C / 8);
  if (nvec == 0) return OGV_OK;
  OGV_DISPATCH_DTYPE(dtype, T, {
    bn_apply_kernel<T><<<flat_grid(nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(x), scale, shift, reinterpret_cast<const T*>(res), reinterpret_cast<T*>(out), nvec,
        C / 8);
    return ogv_check_launch("bn_apply");
  });
}

extern "C" int ogv_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* rstd, float* dgamma,
                                 float* dbeta, long long M, int C, int dtype, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(dy && x && mean && rstd && dgamma && dbeta && C % 8 == 0, "bn_bwd_reduce: bad args");
  OGV_DISPATCH_DTYPE(dtype, T, {
    ColReduceCfg cfg;
    if (!colreduce_config(M, C / 8, &cfg, bn_bwd_reduce_kernel<T>)) { ogv_set_error("bn_bwd_reduce: C=%d too wide", C); return OGV_ERR_UNSUPPORTED; }
    bn_bwd_reduce_kernel<T><<<cfg.grid, cfg.block, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), mean, rstd, dgamma, dbeta, M, C / 8);
    return ogv_check_launch("bn_bwd_reduce");
  });
}

extern "C" int ogv_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* rstd,
                                const float* gamma, const float* dgamma, const float* dbeta, void* dx, long long M,
                                int C, int dtype, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(dy && x && mean && rstd && gamma && dgamma && dbeta && dx && C % 8 == 0, "bn_bwd_apply: bad args");
  const int nv = C / 8;
  if (nv > 256) { ogv_set_error("bn_bwd_apply: C=%d too wide", C); return OGV_ERR_UNSUPPORTED; }
  const int kk = COLREDUCE_THREADS / nv;
  const long long blocks = (M + kk - 1) / kk;
  const long long cap = (long long)ogv_num_sms() * 2;  // 2 x 512 threads x 4 rows x 32 B in flight per SM
  const int grid = (int)(blocks < cap ? blocks : cap);
  OGV_DISPATCH_DTYPE(dtype, T, {
    bn_bwd_apply_kernel<T, 4><<<grid, nv * kk, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), mean, rstd, gamma, dgamma, dbeta,
        reinterpret_cast<T*>(dx), M, nv, 1.f / (float)M);
    return ogv_check_launch("bn_bwd_apply");
  });
}

extern "C" int ogv_bn_act_apply(const void* x, const float* scale, const float* shift, void* out, long long M, int C,
                                int act, int dtype, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(x && scale && shift && out && C > 0 && C % 8 == 0, "bn_act_apply: bad args (C %% 8 == 0)");
  const int nv = C / 8;
  if (nv > 256) { ogv_set_error("bn_act_apply: C=%d too wide", C); return OGV_ERR_UNSUPPORTED; }
  const int kk = COLREDUCE_THREADS / nv;
  const long long blocks = (M + kk - 1) / kk;
  const long long cap = (long long)ogv_num_sms() * 2;
  const int grid = (int)(blocks < cap ? blocks : cap);
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      bn_act_apply_kernel<T, ACT, 4><<<grid, nv * kk, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(x), scale, shift, reinterpret_cast<T*>(out), M, nv);
    });
    return ogv_check_launch("bn_act_apply");
  });
}

extern "C" int ogv_bn_act_bwd_reduce(const void* dy, const void* x, const float* scale, const float* shift,
                                     const float* mean, const float* rstd, float* dgamma, float* dbeta, long long M,
                                     int C, int act, int dtype, void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(dy && x && scale && shift && mean && rstd && dgamma && dbeta && C % 8 == 0, "bn_act_bwd_reduce: bad args");
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      ColReduceCfg cfg;
      if (!colreduce_config(M, C / 8, &cfg, bn_act_bwd_reduce_kernel<T, ACT>)) {
        ogv_set_error("bn_act_bwd_reduce: C=%d too wide", C);
        return OGV_ERR_UNSUPPORTED;
      }
      bn_act_bwd_reduce_kernel<T, ACT><<<cfg.grid, cfg.block, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), scale, shift, mean, rstd, dgamma, dbeta, M,
          C / 8);
    });
    return ogv_check_launch("bn_act_bwd_reduce");
  });
}

extern "C" int ogv_bn_act_bwd_apply(const void* dy, const void* x, const float* scale, const float* shift,
                                    const float* mean, const float* rstd, const float* gamma, const float* dgamma,
                                    const float* dbeta, void* dx, long long M, int C, int act, int dtype,
                                    void* stream) {
  if (M == 0) return OGV_OK;
  OGV_REQUIRE(dy && x && scale && shift && mean && rstd && gamma && dgamma && dbeta && dx && C % 8 == 0,
              "bn_act_bwd_apply: bad args");
  const int nv = C / 8;
  if (nv > 256) { ogv_set_error("bn_act_bwd_apply: C=%d too wide", C); return OGV_ERR_UNSUPPORTED; }
  const int kk = COLREDUCE_THREADS / nv;
  const long long blocks = (M + kk - 1) / kk;
  const long long cap = (long long)ogv_num_sms() * 2;
  const int grid = (int)(blocks < cap ? blocks : cap);
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      bn_act_bwd_apply_kernel<T, ACT, 4><<<grid, nv * kk, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dy), reinterpret_cast<const T*>(x), scale, shift, mean, rstd, gamma, dgamma, dbeta,
          reinterpret_cast<T*>(dx), M, nv, 1.f / (float)M);
    });
    return ogv_check_launch("bn_act_bwd_apply");
  });
}
