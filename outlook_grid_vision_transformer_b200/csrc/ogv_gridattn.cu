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

// =================================================================================================
// Tensor-core path (bf16, N <= 16 tokens per group): the thread-per-row kernels above spend 9*N FMAs per
// element with one shared-memory operand fetch per 4 FMAs -- they are LDS-issue bound, not memory bound.
// Here a WARP owns 16 query rows (= 16/N whole problems) and runs every contraction as
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32) on fragments fetched with ldmatrix: S = Q K^T, O = P V, and
// in backward dP = dO V^T, dQ = dS K, dK = dS^T Q, dV = P^T dO.  Problems smaller than the 16-row tile
// share it block-diagonally (cross-problem scores are masked to -inf).  P and dS are rounded to bf16
// before their second contraction, which is what the reference does under autocast.  (tcgen05 needs
// 64/128-row tiles and TMEM round trips; at N = 4..16 and 0.5-2 % of the model FLOPs the warp-level
// MMA is the right-sized tensor-core instruction.)
// =================================================================================================
constexpr int MA_WARPS = 4;
template <typename K>
int ga_optin(K kernel, size_t bytes);

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// A fragment of the 16 x 16 tile at (0, k0) of a row-major [16][ld] bf16 array
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], const bf16* t, int ld, int k0, int lane) {
  ldsm_x4(a, t + (lane & 15) * ld + k0 + (lane >> 4) * 8);
}
// A fragment of X^T for a row-major [16][ld] X (A[m][k] = X[k][m]), 16 x 16
__device__ __forceinline__ void frag_a_t(uint32_t (&a)[4], const bf16* t, int ld, int lane) {
  ldsm_x4_t(a, t + ((lane & 7) + (lane >> 4) * 8) * ld + ((lane >> 3) & 1) * 8);
}
// B fragments (two n-tiles n0..n0+15) of the k-step k0 for an operand stored [n][k] row-major (e.g. K for Q K^T)
__device__ __forceinline__ void frag_b_nk(uint32_t (&b)[4], const bf16* t, int ld, int n0, int k0, int lane) {
  ldsm_x4(b, t + (n0 + (lane & 7) + (lane >> 4) * 8) * ld + k0 + ((lane >> 3) & 1) * 8);
}
// B fragments (two n-tiles n0..n0+15) of the k-step k0 for an operand stored [k][n] row-major (e.g. V for P V)
__device__ __forceinline__ void frag_b_kn(uint32_t (&b)[4], const bf16* t, int ld, int n0, int k0, int lane) {
  ldsm_x4_t(b, t + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0 + (lane >> 4) * 8);
}

struct MaGeom {
  int B, H, W, C, heads, g, Hg, Wg, N;
  int P16;           // problems per 16-row tile
  long long nprob, ntiles;
  int hd;
  float scale;
};

// Row bookkeeping of a tile, computed ONCE per warp: lane l holds the global row index (or -1 past the
// last problem) and head of tile-row (l & 15); the copy loops fetch them with a shuffle.  Tile rows are
// numbered problem-major: row gr = t*16 + r belongs to problem gr / N, token gr % N (for N < 16 a tile
// holds 16/N whole problems, for N = 16*WP a problem spans WP consecutive tiles).  32-bit arithmetic
// (the host checks that problem and row counts fit).
struct MaRows {
  int m, head;
};
__device__ __forceinline__ MaRows ma_rows(const MaGeom& G, long long t, int lane) {
  const int gr = (int)t * 16 + (lane & 15);
  const int pr = gr / G.N;
  MaRows o;
  o.m = -1;
  o.head = 0;
  if (pr < (int)G.nprob) {
    const int n = gr - pr * G.N;
    o.head = pr % G.heads;
    const int grp = pr / G.heads;
    const int gj = grp % G.g;
    const int gq = grp / G.g;
    const int gi = gq % G.g;
    const int b = gq / G.g;
    o.m = (b * G.H + (n / G.Wg) * G.g + gi) * G.W + (n % G.Wg) * G.g + gj;
  }
  return o;
}

// stage 16 rows x hd columns (zero padded to HDP, zero rows for inactive entries) of a [M, ld] tensor
template <int HDP>
__device__ __forceinline__ void ma_load_tile(bf16* __restrict__ dst, const bf16* __restrict__ src, long long ld,
                                             int col0, const MaGeom& G, const MaRows& R, int lane) {
  constexpr int LD = HDP + 8;
  constexpr int CPR = HDP / 8;  // 16-byte chunks per row
  static_assert((16 * CPR) % 32 == 0, "whole warps of chunks");
#pragma unroll
  for (int i0 = 0; i0 < 16 * CPR; i0 += 32) {
    const int i = i0 + lane;
    const int r = i / CPR, ch = i - r * CPR;
    const int m = __shfl_sync(0xffffffffu, R.m, r);
    const int head = __shfl_sync(0xffffffffu, R.head, r);
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (m >= 0 && ch * 8 < G.hd) u = *reinterpret_cast<const uint4*>(src + (long long)m * ld + col0 + head * G.hd + ch * 8);
    *reinterpret_cast<uint4*>(dst + r * LD + ch * 8) = u;
  }
}
// write 16 rows x hd columns from a staged [16][HDP+8] tile
template <int HDP>
__device__ __forceinline__ void ma_store_tile(const bf16* __restrict__ srcT, bf16* __restrict__ dst, long long ld,
                                              int col0, const MaGeom& G, const MaRows& R, int lane) {
  constexpr int LD = HDP + 8;
  constexpr int CPR = HDP / 8;
#pragma unroll
  for (int i0 = 0; i0 < 16 * CPR; i0 += 32) {
    const int i = i0 + lane;
    const int r = i / CPR, ch = i - r * CPR;
    const int m = __shfl_sync(0xffffffffu, R.m, r);
    const int head = __shfl_sync(0xffffffffu, R.head, r);
    if (m >= 0 && ch * 8 < G.hd)
      *reinterpret_cast<uint4*>(dst + (long long)m * ld + col0 + head * G.hd + ch * 8) =
          *reinterpret_cast<const uint4*>(srcT + r * LD + ch * 8);
  }
}
// C fragments (n-tiles of 8 columns) -> staged bf16 tile
template <int NT>
__device__ __forceinline__ void ma_frags_to_tile(bf16* tile, int ld, const float (&c)[NT][4], int lane) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + g * ld + nt * 8 + q * 2) = pack_bf16(c[nt][0], c[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + (g + 8) * ld + nt * 8 + q * 2) = pack_bf16(c[nt][2], c[nt][3]);
  }
}

// S = scale * Q K^T for this warp's 16 query rows against NKT key tiles (8 keys each), masked
// block-diagonally when several problems share the tile (N < 16), then row softmax in place -> P.
template <int HDP, int NKT>
__device__ __forceinline__ void ma_scores_softmax(float (&s)[NKT][4], const bf16* sQ, const bf16* sK, const MaGeom& G,
                                                  int lane) {
  constexpr int LD = HDP + 8;
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < HDP; k0 += 16) {
    uint32_t a[4];
    frag_a(a, sQ, LD, k0, lane);
#pragma unroll
    for (int n2 = 0; n2 < NKT / 2; ++n2) {
      uint32_t b[4];
      frag_b_nk(b, sK, LD, n2 * 16, k0, lane);
      mma_bf16_16816(s[2 * n2], a, b[0], b[1]);
      mma_bf16_16816(s[2 * n2 + 1], a, b[2], b[3]);
    }
  }
  const int g = lane >> 2, q = lane & 3;
  // rows g (elements 0,1) and g+8 (elements 2,3); columns nt*8 + q*2 + {0,1}
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = g + (i >> 1) * 8, col = nt * 8 + q * 2 + (i & 1);
      const bool same = NKT > 2 || ((row ^ col) & ~(G.N - 1)) == 0;  // N < 16 is a power of two: same block of N
      s[nt][i] = same ? s[nt][i] * G.scale : -INFINITY;
      mx[i >> 1] = fmaxf(mx[i >> 1], s[nt][i]);
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[nt][i] = __expf(s[nt][i] - mx[i >> 1]);  // exp(-inf) = 0 for masked entries
      sum[i >> 1] += s[nt][i];
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
    sum[h] = 1.f / sum[h];
  }
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[nt][i] *= sum[i >> 1];
}

// acc[16 x HDP] += A[16 x 16*KS] * X[16*KS x HDP], A given as accumulator-layout fragments c[2*KS][4] (rounded to
// bf16), X stored row-major [k][n] in shared memory
template <int HDP, int NKT>
__device__ __forceinline__ void ma_mm_frag_kn(float (&acc)[HDP / 8][4], const float (&c)[NKT][4], const bf16* sX,
                                              int lane) {
  constexpr int LD = HDP + 8;
#pragma unroll
  for (int ks = 0; ks < NKT / 2; ++ks) {
    const uint32_t a[4] = {pack_bf16(c[2 * ks][0], c[2 * ks][1]), pack_bf16(c[2 * ks][2], c[2 * ks][3]),
                           pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]), pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3])};
#pragma unroll
    for (int n0 = 0; n0 < HDP; n0 += 16) {
      uint32_t b[4];
      frag_b_kn(b, sX, LD, n0, ks * 16, lane);
      mma_bf16_16816(acc[n0 / 8], a, b[0], b[1]);
      mma_bf16_16816(acc[n0 / 8 + 1], a, b[2], b[3]);
    }
  }
}

// Shared memory of a CTA: 4 tiles of 16 rows each for Q | K | V (| dO), indexed by warp; a problem with
// N = 16*WP tokens owns WP consecutive, WP-aligned tiles, so its K / V rows are contiguous.
template <int HDP, int NKT>
__global__ void __launch_bounds__(MA_WARPS * 32) ma_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                               const MaGeom G) {
  constexpr int LD = HDP + 8;
  constexpr int NTO = HDP / 8;
  constexpr int WP = NKT / 2;  // warps (16-row tiles) per problem
  extern __shared__ __align__(16) uint8_t ma_sm[];
  bf16* const base = reinterpret_cast<bf16*>(ma_sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * MA_WARPS + warp;
  const bool active = t < G.ntiles;
  bf16* sQ = base + (0 * MA_WARPS + warp) * 16 * LD;
  bf16* sK = base + (1 * MA_WARPS + (warp / WP) * WP) * 16 * LD;  // first key row of this warp's problem
  bf16* sV = base + (2 * MA_WARPS + (warp / WP) * WP) * 16 * LD;
  MaRows R;
  R.m = -1; R.head = 0;
  if (active) {
    R = ma_rows(G, t, lane);
    ma_load_tile<HDP>(sQ, qkv, 3LL * G.C, 0, G, R, lane);
    ma_load_tile<HDP>(base + (1 * MA_WARPS + warp) * 16 * LD, qkv, 3LL * G.C, G.C, G, R, lane);
    ma_load_tile<HDP>(base + (2 * MA_WARPS + warp) * 16 * LD, qkv, 3LL * G.C, 2 * G.C, G, R, lane);
  }
  if (WP > 1) __syncthreads(); else __syncwarp();
  if (!active) return;
  float p[NKT][4];
  ma_scores_softmax<HDP, NKT>(p, sQ, sK, G, lane);
  float o[NTO][4];
#pragma unroll
  for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nt][i] = 0.f;
  ma_mm_frag_kn<HDP, NKT>(o, p, sV, lane);  // O = P V
  __syncwarp();  // all lanes are done reading sQ through ldmatrix (only this warp reads its Q tile)
  ma_frags_to_tile<NTO>(sQ, LD, o, lane);
  __syncwarp();
  ma_store_tile<HDP>(sQ, out, G.C, 0, G, R, lane);
}

template <int HDP, int NKT>
__global__ void __launch_bounds__(MA_WARPS * 32) ma_bwd_kernel(const bf16* __restrict__ qkv,
                                                               const bf16* __restrict__ dout,
                                                               bf16* __restrict__ dqkv, const MaGeom G) {
  constexpr int LD = HDP + 8;
  constexpr int NTO = HDP / 8;
  constexpr int WP = NKT / 2;
  constexpr int LDP = NKT * 8 + 8;  // pitch of the P / dS tiles ([16 query rows][keys of the problem])
  extern __shared__ __align__(16) uint8_t ma_sm[];
  bf16* const base = reinterpret_cast<bf16*>(ma_sm);
  bf16* const baseP = base + 4 * MA_WARPS * 16 * LD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * MA_WARPS + warp;
  const bool active = t < G.ntiles;
  const int w0 = (warp / WP) * WP;  // first tile of this warp's problem
  const int sb = warp - w0;         // this warp's 16-row block inside the problem
  bf16* sQ = base + (0 * MA_WARPS + warp) * 16 * LD;
  bf16* sKo = base + (1 * MA_WARPS + warp) * 16 * LD;  // own tiles (loads / output staging)
  bf16* sVo = base + (2 * MA_WARPS + warp) * 16 * LD;
  bf16* sG = base + (3 * MA_WARPS + warp) * 16 * LD;
  bf16* sQp = base + (0 * MA_WARPS + w0) * 16 * LD;    // whole problem (N rows)
  bf16* sKp = base + (1 * MA_WARPS + w0) * 16 * LD;
  bf16* sVp = base + (2 * MA_WARPS + w0) * 16 * LD;
  bf16* sGp = base + (3 * MA_WARPS + w0) * 16 * LD;
  bf16* sP = baseP + (0 * MA_WARPS + warp) * 16 * LDP;  // own query rows
  bf16* sS = baseP + (1 * MA_WARPS + warp) * 16 * LDP;
  bf16* sPp = baseP + (0 * MA_WARPS + w0) * 16 * LDP;   // whole problem
  bf16* sSp = baseP + (1 * MA_WARPS + w0) * 16 * LDP;
  MaRows R;
  R.m = -1; R.head = 0;
  if (active) {
    R = ma_rows(G, t, lane);
    ma_load_tile<HDP>(sQ, qkv, 3LL * G.C, 0, G, R, lane);
    ma_load_tile<HDP>(sKo, qkv, 3LL * G.C, G.C, G, R, lane);
    ma_load_tile<HDP>(sVo, qkv, 3LL * G.C, 2 * G.C, G, R, lane);
    ma_load_tile<HDP>(sG, dout, G.C, 0, G, R, lane);
  }
  if (WP > 1) __syncthreads(); else __syncwarp();
  float acc[NTO][4];
  if (active) {
    // ---- query role: P, dP, dS for this warp's 16 query rows against all keys of the problem ----
    float p[NKT][4];
    ma_scores_softmax<HDP, NKT>(p, sQ, sKp, G, lane);
    float dp[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dp[nt][i] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < HDP; k0 += 16) {  // dP = dO V^T
      uint32_t a[4];
      frag_a(a, sG, LD, k0, lane);
#pragma unroll
      for (int n2 = 0; n2 < NKT / 2; ++n2) {
        uint32_t b[4];
        frag_b_nk(b, sVp, LD, n2 * 16, k0, lane);
        mma_bf16_16816(dp[2 * n2], a, b[0], b[1]);
        mma_bf16_16816(dp[2 * n2 + 1], a, b[2], b[3]);
      }
    }
    // D_i = sum_j P_ij dP_ij ;  dS = P (dP - D) * scale   (masked entries have P = 0)
    float D[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) D[i >> 1] = fmaf(p[nt][i], dp[nt][i], D[i >> 1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      D[h] += __shfl_xor_sync(0xffffffffu, D[h], 1);
      D[h] += __shfl_xor_sync(0xffffffffu, D[h], 2);
    }
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dp[nt][i] = p[nt][i] * (dp[nt][i] - D[i >> 1]) * G.scale;  // dp now holds dS
    ma_frags_to_tile<NKT>(sP, LDP, p, lane);
    ma_frags_to_tile<NKT>(sS, LDP, dp, lane);
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    ma_mm_frag_kn<HDP, NKT>(acc, dp, sKp, lane);  // dQ = dS K
  }
  if (WP > 1) __syncthreads(); else __syncwarp();  // P / dS of every query block of the problem are staged
  float acck[NTO][4], accv[NTO][4];
  if (active) {
    // ---- key role: this warp's 16 keys against all query rows of the problem ----
#pragma unroll
    for (int nt = 0; nt < NTO; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acck[nt][i] = 0.f; accv[nt][i] = 0.f; }
#pragma unroll
    for (int qb = 0; qb < WP; ++qb) {  // query block qb: rows [16 qb, 16 qb + 16) of the problem
      uint32_t at[4], pt[4];
      frag_a_t(at, sSp + qb * 16 * LDP + sb * 16, LDP, lane);  // (dS^T)[own keys][query block]
      frag_a_t(pt, sPp + qb * 16 * LDP + sb * 16, LDP, lane);  // (P^T)
#pragma unroll
      for (int n0 = 0; n0 < HDP; n0 += 16) {
        uint32_t bq[4], bg[4];
        frag_b_kn(bq, sQp, LD, n0, qb * 16, lane);
        frag_b_kn(bg, sGp, LD, n0, qb * 16, lane);
        mma_bf16_16816(acck[n0 / 8], at, bq[0], bq[1]);      // dK = dS^T Q
        mma_bf16_16816(acck[n0 / 8 + 1], at, bq[2], bq[3]);
        mma_bf16_16816(accv[n0 / 8], pt, bg[0], bg[1]);      // dV = P^T dO
        mma_bf16_16816(accv[n0 / 8 + 1], pt, bg[2], bg[3]);
      }
    }
  }
  if (WP > 1) __syncthreads(); else __syncwarp();  // every ldmatrix of the input tiles has completed: reuse them
  if (!active) return;
  ma_frags_to_tile<NTO>(sQ, LD, acc, lane);
  ma_frags_to_tile<NTO>(sKo, LD, acck, lane);
  ma_frags_to_tile<NTO>(sVo, LD, accv, lane);
  __syncwarp();
  ma_store_tile<HDP>(sQ, dqkv, 3LL * G.C, 0, G, R, lane);
  ma_store_tile<HDP>(sKo, dqkv, 3LL * G.C, G.C, G, R, lane);
  ma_store_tile<HDP>(sVo, dqkv, 3LL * G.C, 2 * G.C, G, R, lane);
}

// eligibility: bf16; N divides 16, or N in {32, 64}; head_dim <= 64, a multiple of 8
bool ma_geom(int B, int H, int W, int C, int heads, int g, int dtype, MaGeom* G, int* hdp, int* nkt) {
  if (dtype != OGV_BF16) return false;
  const int Hg = H / g, Wg = W / g, N = Hg * Wg, hd = C / heads;
  if (N < 1) return false;
  if (N <= 16) {
    if ((16 % N) != 0) return false;
    *nkt = 2;
  } else if (N == 32 || N == 64) {
    *nkt = N / 8;
  } else {
    return false;
  }
  if (hd > 64 || (hd % 8) != 0 || (C % 8) != 0) return false;
  G->B = B; G->H = H; G->W = W; G->C = C; G->heads = heads; G->g = g; G->Hg = Hg; G->Wg = Wg; G->N = N;
  G->P16 = N <= 16 ? 16 / N : 1;
  G->nprob = (long long)B * g * g * heads;
  G->ntiles = N <= 16 ? (G->nprob + G->P16 - 1) / G->P16 : G->nprob * (N / 16);
  G->hd = hd;
  G->scale = 1.f / sqrtf((float)hd);
  *hdp = (hd + 15) / 16 * 16;
  // 32-bit row / problem arithmetic in the kernels
  return G->ntiles * 16 + 16 < 0x7fffffffLL && (long long)B * H * W < 0x7fffffffLL;
}

#define MA_DISPATCH_HDP(hdp, ...)                          \
  switch (hdp) {                                           \
    case 16: { constexpr int HDP = 16; __VA_ARGS__; } break; \
    case 32: { constexpr int HDP = 32; __VA_ARGS__; } break; \
    case 48: { constexpr int HDP = 48; __VA_ARGS__; } break; \
    default: { constexpr int HDP = 64; __VA_ARGS__; } break; \
  }
#define MA_DISPATCH_NKT(nkt, ...)                          \
  switch (nkt) {                                           \
    case 2: { constexpr int NKT = 2; __VA_ARGS__; } break; \
    case 4: { constexpr int NKT = 4; __VA_ARGS__; } break; \
    default: { constexpr int NKT = 8; __VA_ARGS__; } break; \
  }

template <int HDP, int NKT>
int ma_launch_fwd(const void* qkv, void* out, const MaGeom& Mg, cudaStream_t st) {
  const size_t smem = (size_t)3 * MA_WARPS * 16 * (HDP + 8) * sizeof(bf16);
  if (int rc = ga_optin(ma_fwd_kernel<HDP, NKT>, smem)) return rc;
  const unsigned grid = (unsigned)((Mg.ntiles + MA_WARPS - 1) / MA_WARPS);
  ma_fwd_kernel<HDP, NKT><<<grid, MA_WARPS * 32, smem, st>>>(reinterpret_cast<const bf16*>(qkv),
                                                            reinterpret_cast<bf16*>(out), Mg);
  return ogv_check_launch("grid_attn_fwd(mma)");
}
template <int HDP, int NKT>
int ma_launch_bwd(const void* qkv, const void* dout, void* dqkv, const MaGeom& Mg, cudaStream_t st) {
  const size_t smem = ((size_t)4 * MA_WARPS * 16 * (HDP + 8) + (size_t)2 * MA_WARPS * 16 * (NKT * 8 + 8)) * sizeof(bf16);
  if (int rc = ga_optin(ma_bwd_kernel<HDP, NKT>, smem)) return rc;
  const unsigned grid = (unsigned)((Mg.ntiles + MA_WARPS - 1) / MA_WARPS);
  ma_bwd_kernel<HDP, NKT><<<grid, MA_WARPS * 32, smem, st>>>(reinterpret_cast<const bf16*>(qkv),
                                                            reinterpret_cast<const bf16*>(dout),
                                                            reinterpret_cast<bf16*>(dqkv), Mg);
  return ogv_check_launch("grid_attn_bwd(mma)");
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
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(qkv && out, "grid_attn_fwd: null pointer");
  GaGeom G;
  int threads;
  if (int rc = ga_geom(B, H, W, C, heads, g, &G, &threads)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  {
    MaGeom Mg;
    int hdp, nkt;
    if (ma_geom(B, H, W, C, heads, g, dtype, &Mg, &hdp, &nkt)) {
      MA_DISPATCH_HDP(hdp, MA_DISPATCH_NKT(nkt, return (ma_launch_fwd<HDP, NKT>(qkv, out, Mg, st))));
    }
  }
  OGV_DISPATCH_DTYPE(dtype, T, GA_DISPATCH_HD(C / heads, return (ga_launch_fwd<T, HD, 0>(qkv, out, nullptr, G, threads, st))));
}

extern "C" int ogv_grid_attn_probs(const void* qkv, float* attn, int B, int H, int W, int C, int heads, int g,
                                   int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(qkv && attn, "grid_attn_probs: null pointer");
  GaGeom G;
  int threads;
  if (int rc = ga_geom(B, H, W, C, heads, g, &G, &threads)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  OGV_DISPATCH_DTYPE(dtype, T, GA_DISPATCH_HD(C / heads, return (ga_launch_fwd<T, HD, 1>(qkv, nullptr, attn, G, threads, st))));
}

extern "C" int ogv_grid_attn_bwd(const void* qkv, const void* dout, void* dqkv, int B, int H, int W, int C,
                                 int heads, int g, int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(qkv && dout && dqkv, "grid_attn_bwd: null pointer");
  GaGeom G;
  int threads;
  if (int rc = ga_geom(B, H, W, C, heads, g, &G, &threads)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  {
    MaGeom Mg;
    int hdp, nkt;
    if (ma_geom(B, H, W, C, heads, g, dtype, &Mg, &hdp, &nkt)) {
      MA_DISPATCH_HDP(hdp, MA_DISPATCH_NKT(nkt, return (ma_launch_bwd<HDP, NKT>(qkv, dout, dqkv, Mg, st))));
    }
  }
  OGV_DISPATCH_DTYPE(dtype, T, GA_DISPATCH_HD(C / heads, return (ga_launch_bwd<T, HD>(qkv, dout, dqkv, G, threads, st))));
}
