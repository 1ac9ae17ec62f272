// Grid attention core on qkv [M, 3C] kept in NHWC row order.  The reference's grid_partition /
// grid_unpartition (grid_partition.py:13-15,29-31) are pure index permutations: group id
// = b*g*g + (h % g)*g + (w % g), token id = (h / g)*Wg + (w / g).  They are folded into the
// addressing here, so no partitioned copy of the activation ever exists.  Per (group, head):
// S = q k^T * hd^-0.5, P = softmax(S), O = P v   (grid_attention.py:70-86); N = Hg*Wg <= 256 tokens,
// so a thread owns one query row and K/V of the group live in shared memory.
#include "ogv_common.cuh"
#include "../../include/ogv.h"

namespace {

struct GaGeom {
  int B, H, W, C, heads, g;
  int Hg, Wg, N;
  long long nprob;  // B*g*g*heads
  int ppc;          // problems per CTA
  float scale;
};

// global row of token n in problem pr; also returns head
__device__ __forceinline__ long long ga_row(const GaGeom& G, long long pr, int n, int& head) {
  head = (int)(pr % G.heads);
  long long grp = pr / G.heads;
  const int gj = (int)(grp % G.g);
  const int gi = (int)((grp / G.g) % G.g);
  const long long b = grp / ((long long)G.g * G.g);
  const int h = (n / G.Wg) * G.g + gi;
  const int w = (n % G.Wg) * G.g + gj;
  return (b * G.H + h) * G.W + w;
}

template <int HD>
struct GaVec {
  static constexpr int V = (HD % 8 == 0) ? 8 : 4;
  static constexpr int LD = HD + 4;  // smem row stride in floats
};

template <typename T, int HD>
__device__ __forceinline__ void ga_load_row(const T* p, float* dst) {
  constexpr int V = GaVec<HD>::V;
#pragma unroll
  for (int d = 0; d < HD; d += V) {
    float v[V];
    ldv<V>(p + d, v);
#pragma unroll
    for (int k = 0; k < V; ++k) dst[d + k] = v[k];
  }
}
template <typename T, int HD>
__device__ __forceinline__ void ga_store_row(T* p, const float* src) {
  constexpr int V = GaVec<HD>::V;
#pragma unroll
  for (int d = 0; d < HD; d += V) {
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = src[d + k];
    stv<V>(p + d, v);
  }
}

template <int HD>
__device__ __forceinline__ float ga_dot(const float (&a)[HD], const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    float4 x = *reinterpret_cast<const float4*>(b + d);
    s = fmaf(a[d], x.x, s);
    s = fmaf(a[d + 1], x.y, s);
    s = fmaf(a[d + 2], x.z, s);
    s = fmaf(a[d + 3], x.w, s);
  }
  return s;
}

// MODE 0: write O.  MODE 1: write the probabilities P (analysis hook).
template <typename T, int HD, int MODE>
__global__ void ga_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ probs, GaGeom G) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LD = GaVec<HD>::LD;
  const int ntok = G.ppc * G.N;
  float* sK = sm;
  float* sV = sm + ntok * LD;
  const int tid = threadIdx.x;
  const int pl = tid / G.N, n = tid % G.N;
  const long long pr = (long long)blockIdx.x * G.ppc + pl;
  const bool active = tid < ntok && pr < G.nprob;
  float q[HD];
  long long m = 0;
  int head = 0;
  if (active) {
    m = ga_row(G, pr, n, head);
    const T* base = qkv + m * 3 * G.C + head * HD;
    ga_load_row<T, HD>(base, q);
    ga_load_row<T, HD>(base + G.C, sK + tid * LD);
    if (MODE == 0) ga_load_row<T, HD>(base + 2 * G.C, sV + tid * LD);
  }
  __syncthreads();
  if (!active) return;
  const float* Kp = sK + pl * G.N * LD;
  const float* Vp = sV + pl * G.N * LD;
  float mx = -INFINITY, l = 0.f;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  for (int j = 0; j < G.N; ++j) {
    const float s = ga_dot<HD>(q, Kp + j * LD) * G.scale;
    const float mn = fmaxf(mx, s);
    const float corr = __expf(mx - mn);
    const float p = __expf(s - mn);
    l = l * corr + p;
    if (MODE == 0) {
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 v = *reinterpret_cast<const float4*>(Vp + j * LD + d);
        o[d] = fmaf(p, v.x, o[d] * corr);
        o[d + 1] = fmaf(p, v.y, o[d + 1] * corr);
        o[d + 2] = fmaf(p, v.z, o[d + 2] * corr);
        o[d + 3] = fmaf(p, v.w, o[d + 3] * corr);
      }
    }
    mx = mn;
  }
  const float inv = 1.f / l;
  if (MODE == 0) {
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= inv;
    ga_store_row<T, HD>(out + m * G.C + head * HD, o);
  } else {
    // probs layout [B*g*g, heads, N, N]; pr = group*heads + head
    float* dst = probs + (pr * G.N + n) * G.N;
    for (int j = 0; j < G.N; ++j) dst[j] = __expf(ga_dot<HD>(q, Kp + j * LD) * G.scale - mx) * inv;
  }
}

template <typename T, int HD>
__global__ void ga_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, T* __restrict__ dqkv, GaGeom G) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LD = GaVec<HD>::LD;
  const int ntok = G.ppc * G.N;
  float* sQ = sm;
  float* sK = sQ + ntok * LD;
  float* sV = sK + ntok * LD;
  float* sG = sV + ntok * LD;   // dO
  float* sM = sG + ntok * LD;   // row max
  float* sL = sM + ntok;        // 1 / row sum
  float* sD = sL + ntok;        // D_i = dO_i . O_i
  const int tid = threadIdx.x;
  const int pl = tid / G.N, n = tid % G.N;
  const long long pr = (long long)blockIdx.x * G.ppc + pl;
  const bool active = tid < ntok && pr < G.nprob;
  long long m = 0;
  int head = 0;
  if (active) {
    m = ga_row(G, pr, n, head);
    const T* base = qkv + m * 3 * G.C + head * HD;
    ga_load_row<T, HD>(base, sQ + tid * LD);
    ga_load_row<T, HD>(base + G.C, sK + tid * LD);
    ga_load_row<T, HD>(base + 2 * G.C, sV + tid * LD);
    ga_load_row<T, HD>(dout + m * G.C + head * HD, sG + tid * LD);
  }
  __syncthreads();
  const float* Qp = sQ + pl * G.N * LD;
  const float* Kp = sK + pl * G.N * LD;
  const float* Vp = sV + pl * G.N * LD;
  const float* Gp = sG + pl * G.N * LD;
  float a[HD], b[HD];  // pass 1/2a: a = q_i, b = dO_i ; pass 2b: a = k_j, b = v_j
  float mx = -INFINITY, l = 0.f;
  if (active) {
#pragma unroll
    for (int d = 0; d < HD; ++d) { a[d] = sQ[tid * LD + d]; b[d] = sG[tid * LD + d]; }
    // pass 1: softmax statistics and D_i = sum_j p_ij (dO_i . v_j)
    float dacc = 0.f;
    for (int j = 0; j < G.N; ++j) {
      const float s = ga_dot<HD>(a, Kp + j * LD) * G.scale;
      const float mn = fmaxf(mx, s);
      const float corr = __expf(mx - mn);
      const float p = __expf(s - mn);
      l = l * corr + p;
      dacc = dacc * corr + p * ga_dot<HD>(b, Vp + j * LD);
      mx = mn;
    }
    const float inv = 1.f / l;
    sM[tid] = mx;
    sL[tid] = inv;
    sD[tid] = dacc * inv;
  }
  __syncthreads();
  if (!active) return;
  T* gbase = dqkv + m * 3 * G.C + head * HD;
  {
    // pass 2a (query role): dq_i = scale * sum_j dS_ij k_j
    const float inv = sL[tid], Di = sD[tid];
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    for (int j = 0; j < G.N; ++j) {
      const float s = ga_dot<HD>(a, Kp + j * LD) * G.scale;
      const float p = __expf(s - mx) * inv;
      const float dS = p * (ga_dot<HD>(b, Vp + j * LD) - Di) * G.scale;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 k = *reinterpret_cast<const float4*>(Kp + j * LD + d);
        dq[d] = fmaf(dS, k.x, dq[d]);
        dq[d + 1] = fmaf(dS, k.y, dq[d + 1]);
        dq[d + 2] = fmaf(dS, k.z, dq[d + 2]);
        dq[d + 3] = fmaf(dS, k.w, dq[d + 3]);
      }
    }
    ga_store_row<T, HD>(gbase, dq);
  }
  {
    // pass 2b (key role): dk_j = scale * sum_i dS_ij q_i ; dv_j = sum_i p_ij dO_i
#pragma unroll
    for (int d = 0; d < HD; ++d) { a[d] = sK[tid * LD + d]; b[d] = sV[tid * LD + d]; }
    float dk[HD], dv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    const float* Mp = sM + pl * G.N;
    const float* Lp = sL + pl * G.N;
    const float* Dp = sD + pl * G.N;
    for (int i = 0; i < G.N; ++i) {
      const float s = ga_dot<HD>(a, Qp + i * LD) * G.scale;
      const float p = __expf(s - Mp[i]) * Lp[i];
      const float dS = p * (ga_dot<HD>(b, Gp + i * LD) - Dp[i]) * G.scale;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 qv = *reinterpret_cast<const float4*>(Qp + i * LD + d);
        float4 gv = *reinterpret_cast<const float4*>(Gp + i * LD + d);
        dk[d] = fmaf(dS, qv.x, dk[d]);
        dk[d + 1] = fmaf(dS, qv.y, dk[d + 1]);
        dk[d + 2] = fmaf(dS, qv.z, dk[d + 2]);
        dk[d + 3] = fmaf(dS, qv.w, dk[d + 3]);
        dv[d] = fmaf(p, gv.x, dv[d]);
        dv[d + 1] = fmaf(p, gv.y, dv[d + 1]);
        dv[d + 2] = fmaf(p, gv.z, dv[d + 2]);
        dv[d + 3] = fmaf(p, gv.w, dv[d + 3]);
      }
    }
    ga_store_row<T, HD>(gbase + G.C, dk);
    ga_store_row<T, HD>(gbase + 2 * G.C, dv);
  }
}

int ga_geom(int B, int H, int W, int C, int heads, int g, GaGeom* G, int* threads) {
  if (g <= 0 || heads <= 0 || C <= 0 || H <= 0 || W <= 0) { ogv_set_error("grid_attn: non-positive dims"); return OGV_ERR_ARG; }
  if (H % g || W % g) { ogv_set_error("grid_attn: H and W must be divisible by grid_size (H=%d W=%d g=%d)", H, W, g); return OGV_ERR_ARG; }
  if (C % heads) { ogv_set_error("grid_attn: dim (%d) must be divisible by num_heads (%d)", C, heads); return OGV_ERR_ARG; }
  G->B = B; G->H = H; G->W = W; G->C = C; G->heads = heads; G->g = g;
  G->Hg = H / g; G->Wg = W / g; G->N = G->Hg * G->Wg;
  if (G->N > 256) { ogv_set_error("grid_attn: %d tokens per group > 256 unsupported", G->N); return OGV_ERR_UNSUPPORTED; }
  G->nprob = (long long)B * g * g * heads;
  *threads = G->N <= 128 ? 128 : 256;
  G->ppc = *threads / G->N;
  G->scale = 1.f / sqrtf((float)(C / heads));
  return OGV_OK;
}

template <typename K>
int ga_optin(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) { ogv_set_error("grid_attn: %zu B of shared memory needed", bytes); return OGV_ERR_UNSUPPORTED; }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { ogv_set_error("grid_attn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return OGV_ERR_CUDA; }
  return OGV_OK;
}

template <typename T, int HD, int MODE>
int ga_launch_fwd(const void* qkv, void* out, float* probs, const GaGeom& G, int threads, cudaStream_t st) {
  size_t bytes = (size_t)2 * G.ppc * G.N * GaVec<HD>::LD * sizeof(float);
  if (int rc = ga_optin(ga_fwd_kernel<T, HD, MODE>, bytes)) return rc;
  long long grid = (G.nprob + G.ppc - 1) / G.ppc;
  ga_fwd_kernel<T, HD, MODE><<<(unsigned)grid, threads, bytes, st>>>(reinterpret_cast<const T*>(qkv),
                                                                      reinterpret_cast<T*>(out), probs, G);
  return ogv_check_launch("grid_attn_fwd");
}
template <typename T, int HD>
int ga_launch_bwd(const void* qkv, const void* dout, void* dqkv, const GaGeom& G, int threads, cudaStream_t st) {
  size_t ntok = (size_t)G.ppc * G.N;
  size_t bytes = (4 * ntok * GaVec<HD>::LD + 3 * ntok) * sizeof(float);
  if (int rc = ga_optin(ga_bwd_kernel<T, HD>, bytes)) return rc;
  long long grid = (G.nprob + G.ppc - 1) / G.ppc;
  ga_bwd_kernel<T, HD><<<(unsigned)grid, threads, bytes, st>>>(reinterpret_cast<const T*>(qkv),
                                                               reinterpret_cast<const T*>(dout),
                                                               reinterpret_cast<T*>(dqkv), G);
  return ogv_check_launch("grid_attn_bwd");
}

#define GA_DISPATCH_HD(hd, ...)                                                                  \
  switch (hd) {                                                                                  \
    case 4: { constexpr int HD = 4; __VA_ARGS__; }                                               \
    case 8: { constexpr int HD = 8; __VA_ARGS__; }                                               \
    case 16: { constexpr int HD = 16; __VA_ARGS__; }                                             \
    case 24: { constexpr int HD = 24; __VA_ARGS__; }                                             \
    case 32: { constexpr int HD = 32; __VA_ARGS__; }                                             \
    case 40: { constexpr int HD = 40; __VA_ARGS__; }                                             \
    case 56: { constexpr int HD = 56; __VA_ARGS__; }                                             \
    case 48: { constexpr int HD = 48; __VA_ARGS__; }                                             \
    case 64: { constexpr int HD = 64; __VA_ARGS__; }                                             \
    default:                                                                                     \
      ogv_set_error("grid_attn: head_dim %d not in {4,8,16,24,32,40,48,56,64}", hd);                   \
      return OGV_ERR_UNSUPPORTED;                                                                \
  }

}  // namespace

extern "C" int ogv_grid_attn_fwd(const void* qkv, void* out, int B, int H, int W, int C, int heads, int g,
                                 int dtype, void* stream) {
  OGV_REQUIRE(qkv && out, "grid_attn_fwd: null pointer");
  GaGeom G;
  int threads;
  if (int rc = ga_geom(B, H, W, C, heads, g, &G, &threads)) return rc;
  if (B == 0) return OGV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, GA_DISPATCH_HD(C / heads, return (ga_launch_fwd<T, HD, 0>(qkv, out, nullptr, G, threads, st))));
}

extern "C" int ogv_grid_attn_probs(const void* qkv, float* attn, int B, int H, int W, int C, int heads, int g,
                                   int dtype, void* stream) {
  OGV_REQUIRE(qkv && attn, "grid_attn_probs: null pointer");
  GaGeom G;
  int threads;
  if (int rc = ga_geom(B, H, W, C, heads, g, &G, &threads)) return rc;
  if (B == 0) return OGV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, GA_DISPATCH_HD(C / heads, return (ga_launch_fwd<T, HD, 1>(qkv, nullptr, attn, G, threads, st))));
}

extern "C" int ogv_grid_attn_bwd(const void* qkv, const void* dout, void* dqkv, int B, int H, int W, int C,
                                 int heads, int g, int dtype, void* stream) {
  OGV_REQUIRE(qkv && dout && dqkv, "grid_attn_bwd: null pointer");
  GaGeom G;
  int threads;
  if (int rc = ga_geom(B, H, W, C, heads, g, &G, &threads)) return rc;
  if (B == 0) return OGV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, GA_DISPATCH_HD(C / heads, return (ga_launch_bwd<T, HD>(qkv, dout, dqkv, G, threads, st))));
}
