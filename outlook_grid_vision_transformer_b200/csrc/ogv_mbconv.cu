// MBConv interior around the depthwise stencil (ogv_dwconv.cu): SqueezeExcite pooling / gating and the
// BatchNorm2 + activation + gate backward passes.
// Reference: mbc_conv.py:9-27 (SqueezeExcite), :44-98 (MBConv); math in SURVEY Appendix A.3/A.4.
//
// All kernels here stream [B, HW, Cm] tensors image by image: a CTA is 32 channel-vectors (8
// channels each, 16-byte accesses, 512 contiguous bytes per warp row) x 8 row lanes; every
// per-channel / per-(image, channel) parameter is loaded once into registers, so the inner loop is
// nothing but the tensor traffic.  No integer division anywhere on the data path.
#include "ogv_common.cuh"
#include "../../include/ogv.h"

namespace {

constexpr int IMG_THREADS = 256;
constexpr int IMG_ROWLANES = 8;
// rows in flight per thread in the two-operand streaming passes (A/B-tunable)
#ifndef MBS_U
#define MBS_U 2
#endif
#ifndef DBA_U
#define DBA_U 4
#endif

// u = sc*x + sh
#define OGV_BN_U(k) fmaf(x[k], sc[k], sh[k])

// ------------------------------------------------------------------------------------------------
// SE squeeze (MODE 0) and its backward reduction (MODE 1), one (image, 32-vector block) per CTA
// MODE 0: out[b,c] = (1/HW) sum_p act(sc*x+sh)      MODE 1: out[b,c] = sum_p y[p,c] * act(sc*x+sh)
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, int ACT>
__global__ void __launch_bounds__(IMG_THREADS) se_reduce_kernel(const T* __restrict__ xin, const T* __restrict__ yin,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                float* __restrict__ out, int HW, int Cm) {
  __shared__ float red[IMG_ROWLANES][32][9];
  const int nv = Cm / 8;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int cv = blockIdx.y * 32 + cx;
  const int b = blockIdx.x;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (cv < nv) {
    float sc[8], sh[8];
    ld8(scale + cv * 8, sc);
    ld8(shift + cv * 8, sh);
    const T* xp = xin + ((long long)b * HW) * Cm + cv * 8;
    const T* yp = MODE == 1 ? yin + ((long long)b * HW) * Cm + cv * 8 : nullptr;
#pragma unroll 4
    for (int p = ry; p < HW; p += IMG_ROWLANES) {
      float x[8];
      ld8(xp + (long long)p * Cm, x);
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += act_apply_t<ACT, FastAct<T>::value>(OGV_BN_U(k));
      } else {
        float yv[8];
        ld8(yp + (long long)p * Cm, yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(yv[k], act_apply_t<ACT, FastAct<T>::value>(OGV_BN_U(k)), acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[ry][cx][k] = acc[k];
  __syncthreads();
  if (ry == 0 && cv < nv) {
    const float norm = MODE == 0 ? 1.f / (float)HW : 1.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < IMG_ROWLANES; ++j) s += red[j][cx][k];
      out[(long long)b * Cm + cv * 8 + k] = s * norm;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One pass over (dd_act, d_pre) producing every per-(image, channel) sum the SE backward AND the
// BatchNorm2 backward reductions need (a = act(u), a' = act'(u), u = sc*x + sh, xh = (x-mean)*rstd,
// g = dd_act):   S0 = sum g*a   S1 = sum g*a'   S2 = sum g*a'*xh   S3 = sum a'   S4 = sum a'*xh
// With du = (g*gate + dpool/HW) * a' (mbc_conv.py:27 backward):
//   dgate[b,c] = S0 ;  dbeta2[c] = sum_b gate*S1 + dpool/HW*S3 ;  dgamma2[c] = sum_b gate*S2 + dpool/HW*S4
// so the wide tensors are read once instead of twice.   stats layout: [5][B][Cm] fp32.
// ------------------------------------------------------------------------------------------------
template <typename T, int ACT>
__global__ void __launch_bounds__(IMG_THREADS) mbconv_bwd_stats_kernel(
    const T* __restrict__ dd_act, const T* __restrict__ d_pre, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ stats, int B, int HW, int Cm) {
  __shared__ float red[IMG_ROWLANES][32][9];
  const int nv = Cm / 8;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int cv = blockIdx.y * 32 + cx;
  const int b = blockIdx.x;
  // The loop accumulates the xh-weighted sums against x itself (A2 = sum g*a'*x, A4 = sum a'*x); the last step turns
  // them into S2 = rstd*(A2 - mean*S1), S4 = rstd*(A4 - mean*S3).  That keeps mean / rstd out of the loop's registers,
  // which pays for a second set of row registers: the loads of batch i+1 are in flight while batch i is being reduced
  // (with 4 warps per scheduler and ~300 issue slots of math per batch, load -> math -> load left the SM idle for
  // most of every DRAM round trip: 48 % of HBM before, measured).
  float acc[5][8];
#pragma unroll
  for (int q = 0; q < 5; ++q)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
  if (cv < nv) {
    float sc[8], sh[8];
    ld8(scale + cv * 8, sc);
    ld8(shift + cv * 8, sh);
    const T* xp = d_pre + ((long long)b * HW) * Cm + cv * 8;
    const T* gp = dd_act + ((long long)b * HW) * Cm + cv * 8;
    constexpr int U = MBS_U;  // rows per batch
    constexpr int STEP = IMG_ROWLANES * U;
    Raw8<T> rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = ry + u * IMG_ROWLANES;
      if (p < HW) {
        ld_raw8(xp + (long long)p * Cm, rx[u]);
        ld_raw8(gp + (long long)p * Cm, rg[u]);
      }
    }
    for (int p0 = ry; p0 < HW; p0 += STEP) {
      Raw8<T> nx[U], ng[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + STEP + u * IMG_ROWLANES;
        if (p < HW) {
          ld_raw8(xp + (long long)p * Cm, nx[u]);
          ld_raw8(gp + (long long)p * Cm, ng[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p0 + u * IMG_ROWLANES < HW) {
          float x[8], g[8];
          cvt_raw8(rx[u], x);
          cvt_raw8(rg[u], g);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float uu = OGV_BN_U(k);
            float a, da;
            act_both_t<ACT, FastAct<T>::value>(uu, &a, &da);
            const float gda = g[k] * da;
            acc[0][k] = fmaf(g[k], a, acc[0][k]);
            acc[1][k] += gda;
            acc[2][k] = fmaf(gda, x[k], acc[2][k]);
            acc[3][k] += da;
            acc[4][k] = fmaf(da, x[k], acc[4][k]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rx[u] = nx[u];
        rg[u] = ng[u];
      }
    }
  }
  const long long plane = (long long)B * Cm;
  float tot[5][8];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[ry][cx][k] = acc[q][k];
    __syncthreads();
    if (ry == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < IMG_ROWLANES; ++j) sum += red[j][cx][k];
        tot[q][k] = sum;
      }
    }
    __syncthreads();
  }
  if (ry == 0 && cv < nv) {
    float mu[8], rs[8];
    ld8(mean + cv * 8, mu);
    ld8(rstd + cv * 8, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      tot[2][k] = rs[k] * fmaf(-mu[k], tot[1][k], tot[2][k]);
      tot[4][k] = rs[k] * fmaf(-mu[k], tot[3][k], tot[4][k]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float* dst = stats + q * plane + (long long)b * Cm + cv * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(tot[q][0], tot[q][1], tot[q][2], tot[q][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(tot[q][4], tot[q][5], tot[q][6], tot[q][7]);
    }
  }
}

// dbeta2[c] += sum_b gate*S1 + dpool/HW*S3 ;  dgamma2[c] += sum_b gate*S2 + dpool/HW*S4
__global__ void __launch_bounds__(1024) mbconv_bn2_finalize_kernel(const float* __restrict__ stats,
                                                                   const float* __restrict__ gate,
                                                                   const float* __restrict__ dpool,
                                                                   float* __restrict__ dgamma2,
                                                                   float* __restrict__ dbeta2, int B, int Cm,
                                                                   float inv_hw) {
  __shared__ float rb[32][33], rg[32][33];
  const int cx = threadIdx.x, by = threadIdx.y;
  const int c = blockIdx.x * 32 + cx;
  const long long plane = (long long)B * Cm;
  float db = 0.f, dg = 0.f;
  if (c < Cm) {
#pragma unroll 4
    for (int b = by; b < B; b += 32) {
      const long long o = (long long)b * Cm + c;
      const float gt = gate[o], dp = dpool[o] * inv_hw;
      db += gt * stats[1 * plane + o] + dp * stats[3 * plane + o];
      dg += gt * stats[2 * plane + o] + dp * stats[4 * plane + o];
    }
  }
  rb[by][cx] = db;
  rg[by][cx] = dg;
  __syncthreads();
  if (by == 0 && c < Cm) {
    float sb = 0.f, sg = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { sb += rb[j][cx]; sg += rg[j][cx]; }
    dbeta2[c] += sb;
    dgamma2[c] += sg;
  }
}

// ------------------------------------------------------------------------------------------------
// elementwise, image by image: grid (B, ceil(nv/32), row splits)
// ------------------------------------------------------------------------------------------------
// d_act = act(sc*x + sh) * gate[b,c]
template <typename T, int ACT>
__global__ void __launch_bounds__(IMG_THREADS) bn_act_gate_kernel(const T* __restrict__ xin,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift,
                                                                  const float* __restrict__ gate, T* __restrict__ out,
                                                                  int HW, int Cm) {
  const int nv = Cm / 8;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int cv = blockIdx.y * 32 + cx;
  const int b = blockIdx.x;
  if (cv >= nv) return;
  float sc[8], sh[8], gt[8];
  ld8(scale + cv * 8, sc);
  ld8(shift + cv * 8, sh);
  ld8(gate + (long long)b * Cm + cv * 8, gt);
  const long long base = ((long long)b * HW) * Cm + cv * 8;
  const int step = IMG_ROWLANES * gridDim.z;
#pragma unroll 4
  for (int p = blockIdx.z * IMG_ROWLANES + ry; p < HW; p += step) {
    float x[8];
    ld8(xin + base + (long long)p * Cm, x);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = act_apply_t<ACT, FastAct<T>::value>(OGV_BN_U(k)) * gt[k];
    st8(out + base + (long long)p * Cm, x);
  }
}

// du = (g*gate + dpool/HW) * act'(u) ;  dd_pre = gamma2*rstd2*(du - dbeta2/n - xh*dgamma2/n)
template <typename T, int ACT>
__global__ void __launch_bounds__(IMG_THREADS) dw_bn2_bwd_apply_kernel(
    const T* __restrict__ dd_act, const T* __restrict__ d_pre, const float* __restrict__ gate,
    const float* __restrict__ dpool, const float* __restrict__ scale2, const float* __restrict__ shift2,
    const float* __restrict__ mean2, const float* __restrict__ rstd2, const float* __restrict__ gamma2,
    const float* __restrict__ dgamma2, const float* __restrict__ dbeta2, T* __restrict__ dd_pre, int HW, int Cm,
    float inv_n) {
  const int nv = Cm / 8;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int cv = blockIdx.y * 32 + cx;
  const int b = blockIdx.x;
  if (cv >= nv) return;
  // per-channel constants folded so the inner loop is:  o = A*du - Bc - Cx*x   (xh = (x-mu)*rs)
  float sc[8], sh[8], gt[8], dp[8], A[8], Bc[8], Cx[8];
  {
    float mu[8], rs[8], gm[8], dg[8], db[8];
    ld8(scale2 + cv * 8, sc);
    ld8(shift2 + cv * 8, sh);
    ld8(mean2 + cv * 8, mu);
    ld8(rstd2 + cv * 8, rs);
    ld8(gamma2 + cv * 8, gm);
    ld8(dgamma2 + cv * 8, dg);
    ld8(dbeta2 + cv * 8, db);
    ld8(gate + (long long)b * Cm + cv * 8, gt);
    ld8(dpool + (long long)b * Cm + cv * 8, dp);
    const float inv_hw = 1.f / (float)HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      dp[k] *= inv_hw;
      A[k] = gm[k] * rs[k];
      Cx[k] = A[k] * rs[k] * dg[k] * inv_n;               // coefficient of x
      Bc[k] = A[k] * db[k] * inv_n - Cx[k] * mu[k];       // constant part
    }
  }
  const long long base = ((long long)b * HW) * Cm + cv * 8;
  const int step = IMG_ROWLANES * gridDim.z;
  constexpr int U = DBA_U;  // rows in flight per thread
  for (int p0 = blockIdx.z * IMG_ROWLANES + ry; p0 < HW; p0 += step * U) {
    Raw8<T> rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * step;
      if (p < HW) {
        ld_raw8(d_pre + base + (long long)p * Cm, rx[u]);
        ld_raw8(dd_act + base + (long long)p * Cm, rg[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * step;
      if (p < HW) {
        float x[8], g[8], o[8];
        cvt_raw8(rx[u], x);
        cvt_raw8(rg[u], g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float du = fmaf(g[k], gt[k], dp[k]) * act_grad_t<ACT, FastAct<T>::value>(OGV_BN_U(k));
          o[k] = fmaf(A[k], du, -fmaf(Cx[k], x[k], Bc[k]));
        }
        st8(dd_pre + base + (long long)p * Cm, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Squeeze-excite MLP on [B, Cm] rows as ONE kernel per direction (bf16 compute mode, stages whose two weight matrices
// are small: Cs a power of two <= 256, Cm a multiple of 256).  The two 1x1 convs (mbc_conv.py:17-19,24-26) are a few
// MFLOP per block: as tcgen05 GEMM launches they cost 10-25 us EACH (8-CTA grids, pure latency), plus the casts, the
// sigma' pass and the bias-gradient sums around them -- 13 launches per block and step.  Here a CTA owns RB = 8 images,
// keeps their rows in shared memory ([k][image], so one broadcast LDS.128 pair feeds 8 FMAs) and runs both layers with
// the K-major weight copy streamed from L2 once per CTA; operands are rounded to bf16 exactly where the GEMM route
// rounds them (inputs, the hidden activation), accumulation in fp32.
//   forward : s1_pre = pool W1^T + b1, s1a = act(s1_pre), gate_pre = s1a W2^T + b2, gate = sigmoid(gate_pre)
//   backward: dgp = dgate * sigmoid'(gate_pre), ds1_pre = (dgp W2) * act'(s1_pre), dpool = ds1_pre W1
// ------------------------------------------------------------------------------------------------
constexpr int SE_RB = 8;
constexpr int SE_THREADS = 1024;  // 32 warps: the loops below are chains of L2 loads, latency is hidden by warps
constexpr int SE_TN = 256;        // output columns per round; SE_THREADS / SE_TN k-groups share a column

// out[r][n] = sum_k in_s[k][r] * Wt[k][n] for n in [0, N), all RB images; epi(n, acc) is called by ONE thread per n.
template <typename Epi>
__device__ __forceinline__ void se_layer(const float* __restrict__ in_s, const bf16* __restrict__ Wt, int K, int N,
                                         float* __restrict__ red_s, Epi&& epi) {
  const int tid = threadIdx.x;
  const int TN = N < SE_TN ? N : SE_TN;             // N is a power of two <= 256 or a multiple of 256
  const int G = SE_THREADS / TN;                    // k-groups sharing an output column (>= 4)
  const int tn = tid % TN, kg = tid / TN;
  const int kper = K / G;
  for (int n0 = 0; n0 < N; n0 += TN) {
    const int n = n0 + tn;
    float acc[SE_RB];
#pragma unroll
    for (int r = 0; r < SE_RB; ++r) acc[r] = 0.f;
    const bf16* wp = Wt + (long long)(kg * kper) * N + n;
    const float* ip = in_s + (kg * kper) * SE_RB;
#pragma unroll 8
    for (int k = 0; k < kper; ++k) {
      const float w = __bfloat162float(wp[(long long)k * N]);
      const float4 a0 = *reinterpret_cast<const float4*>(ip + k * SE_RB);
      const float4 a1 = *reinterpret_cast<const float4*>(ip + k * SE_RB + 4);
      acc[0] = fmaf(w, a0.x, acc[0]); acc[1] = fmaf(w, a0.y, acc[1]);
      acc[2] = fmaf(w, a0.z, acc[2]); acc[3] = fmaf(w, a0.w, acc[3]);
      acc[4] = fmaf(w, a1.x, acc[4]); acc[5] = fmaf(w, a1.y, acc[5]);
      acc[6] = fmaf(w, a1.z, acc[6]); acc[7] = fmaf(w, a1.w, acc[7]);
    }
    if (G > 1) {
      __syncthreads();  // red_s may still be read by the previous round
#pragma unroll
      for (int r = 0; r < SE_RB; ++r) red_s[(kg * TN + tn) * SE_RB + r] = acc[r];
      __syncthreads();
      if (kg == 0) {
        for (int gdx = 1; gdx < G; ++gdx)
#pragma unroll
          for (int r = 0; r < SE_RB; ++r) acc[r] += red_s[(gdx * TN + tn) * SE_RB + r];
      }
    }
    if (kg == 0) epi(n, acc);
  }
}

template <int ACT>
__global__ void __launch_bounds__(SE_THREADS)
se_mlp_fwd_kernel(const float* __restrict__ pool, const bf16* __restrict__ w1t, const float* __restrict__ b1,
                  const bf16* __restrict__ w2t, const float* __restrict__ b2, bf16* __restrict__ pool_c,
                  bf16* __restrict__ s1_pre, bf16* __restrict__ s1a, float* __restrict__ gate_pre,
                  float* __restrict__ gate, int B, int Cm, int Cs) {
  extern __shared__ __align__(16) float se_sm[];
  float* const in_s = se_sm;                       // [Cm][RB]
  float* const h_s = se_sm + Cm * SE_RB;           // [Cs][RB]
  float* const red_s = h_s + Cs * SE_RB;           // [SE_THREADS][RB]
  const int b0 = blockIdx.x * SE_RB;
  for (int i = threadIdx.x; i < Cm * SE_RB; i += SE_THREADS) {
    const int r = i / Cm, k = i - r * Cm;          // coalesced over k
    float v = 0.f;
    if (b0 + r < B) {
      v = round_to<bf16>(pool[(long long)(b0 + r) * Cm + k]);
      pool_c[(long long)(b0 + r) * Cm + k] = __float2bfloat16_rn(v);
    }
    in_s[k * SE_RB + r] = v;
  }
  __syncthreads();
  se_layer(in_s, w1t, Cm, Cs, red_s, [&](int n, const float (&acc)[SE_RB]) {
    const float bias = b1[n];
#pragma unroll
    for (int r = 0; r < SE_RB; ++r) {
      const float pre = round_to<bf16>(acc[r] + bias);
      const float a = round_to<bf16>(act_apply_t<ACT, true>(acc[r] + bias));
      h_s[n * SE_RB + r] = a;
      if (b0 + r < B) {
        s1_pre[(long long)(b0 + r) * Cs + n] = __float2bfloat16_rn(pre);
        s1a[(long long)(b0 + r) * Cs + n] = __float2bfloat16_rn(a);
      }
    }
  });
  __syncthreads();
  se_layer(h_s, w2t, Cs, Cm, red_s, [&](int n, const float (&acc)[SE_RB]) {
    const float bias = b2[n];
#pragma unroll
    for (int r = 0; r < SE_RB; ++r) {
      if (b0 + r < B) {
        const float pre = acc[r] + bias;
        gate_pre[(long long)(b0 + r) * Cm + n] = pre;
        gate[(long long)(b0 + r) * Cm + n] = sigmoid_f(pre);
      }
    }
  });
}

template <int ACT>
__global__ void __launch_bounds__(SE_THREADS)
se_mlp_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate_pre, const bf16* __restrict__ s1_pre,
                  const bf16* __restrict__ w2, const bf16* __restrict__ w1, bf16* __restrict__ dgate_c,
                  bf16* __restrict__ ds1_pre, float* __restrict__ dpool, int B, int Cm, int Cs) {
  extern __shared__ __align__(16) float se_sm[];
  float* const in_s = se_sm;                       // [Cm][RB]
  float* const h_s = se_sm + Cm * SE_RB;           // [Cs][RB]
  float* const red_s = h_s + Cs * SE_RB;
  const int b0 = blockIdx.x * SE_RB;
  for (int i = threadIdx.x; i < Cm * SE_RB; i += SE_THREADS) {
    const int r = i / Cm, k = i - r * Cm;
    float v = 0.f;
    if (b0 + r < B) {
      const long long o = (long long)(b0 + r) * Cm + k;
      const float sg = sigmoid_f(gate_pre[o]);
      v = round_to<bf16>(dgate[o] * sg * (1.f - sg));
      dgate_c[o] = __float2bfloat16_rn(v);
    }
    in_s[k * SE_RB + r] = v;
  }
  __syncthreads();
  se_layer(in_s, w2, Cm, Cs, red_s, [&](int n, const float (&acc)[SE_RB]) {
#pragma unroll
    for (int r = 0; r < SE_RB; ++r) {
      float d = 0.f;
      if (b0 + r < B) {
        const long long o = (long long)(b0 + r) * Cs + n;
        d = round_to<bf16>(acc[r] * act_grad_t<ACT, false>(__bfloat162float(s1_pre[o])));
        ds1_pre[o] = __float2bfloat16_rn(d);
      }
      h_s[n * SE_RB + r] = d;
    }
  });
  __syncthreads();
  se_layer(h_s, w1, Cs, Cm, red_s, [&](int n, const float (&acc)[SE_RB]) {
#pragma unroll
    for (int r = 0; r < SE_RB; ++r)
      if (b0 + r < B) dpool[(long long)(b0 + r) * Cm + n] = acc[r];
  });
}

inline bool se_mlp_shape_ok(int Cm, int Cs) {
  // k-groups must divide both reduction lengths: K / G with G = SE_THREADS / min(N, SE_TN)
  const bool cs_ok = Cs >= 16 && Cs <= 256 && (Cs & (Cs - 1)) == 0;
  if (!(cs_ok && Cm >= 256 && Cm % 256 == 0 && Cm <= 2048)) return false;
  const int g1 = SE_THREADS / (Cs < SE_TN ? Cs : SE_TN), g2 = SE_THREADS / SE_TN;
  return Cm % g1 == 0 && Cs % g2 == 0;
}
inline size_t se_mlp_smem(int Cm, int Cs) { return (size_t)(Cm + Cs + SE_THREADS) * SE_RB * sizeof(float); }

// row splits so that small batches still fill the machine
inline int img_row_splits(int B, int nvb, int HW) {
  long long ctas = (long long)B * nvb;
  long long want = ((long long)ogv_num_sms() * 8 + ctas - 1) / ctas;
  long long maxs = (HW + IMG_ROWLANES * 4 - 1) / (IMG_ROWLANES * 4);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

}  // namespace

extern "C" int ogv_se_pool(const void* d_pre, const float* scale2, const float* shift2, float* pool, int B, int HW,
                           int Cm, int act, int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(d_pre && scale2 && shift2 && pool && Cm % 8 == 0 && HW > 0, "se_pool: bad args");
  dim3 grid(B, ogv_ceil_div(Cm / 8, 32));
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      se_reduce_kernel<T, 0, ACT><<<grid, IMG_THREADS, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(d_pre), nullptr, scale2, shift2, pool, HW, Cm);
    });
    return ogv_check_launch("se_pool");
  });
}

extern "C" int ogv_se_bwd_reduce(const void* dd_act, const void* d_pre, const float* scale2, const float* shift2,
                                 float* dgate, int B, int HW, int Cm, int act, int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(dd_act && d_pre && scale2 && shift2 && dgate && Cm % 8 == 0 && HW > 0, "se_bwd_reduce: bad args");
  dim3 grid(B, ogv_ceil_div(Cm / 8, 32));
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      se_reduce_kernel<T, 1, ACT><<<grid, IMG_THREADS, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(d_pre), reinterpret_cast<const T*>(dd_act), scale2, shift2, dgate, HW, Cm);
    });
    return ogv_check_launch("se_bwd_reduce");
  });
}

extern "C" int ogv_mbconv_bwd_stats(const void* dd_act, const void* d_pre, const float* scale2, const float* shift2,
                                    const float* mean2, const float* rstd2, float* stats, int B, int HW, int Cm,
                                    int act, int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(dd_act && d_pre && scale2 && shift2 && mean2 && rstd2 && stats && Cm % 8 == 0 && HW > 0,
              "mbconv_bwd_stats: bad args");
  dim3 grid(B, ogv_ceil_div(Cm / 8, 32));
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      mbconv_bwd_stats_kernel<T, ACT><<<grid, IMG_THREADS, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dd_act), reinterpret_cast<const T*>(d_pre), scale2, shift2, mean2, rstd2, stats,
          B, HW, Cm);
    });
    return ogv_check_launch("mbconv_bwd_stats");
  });
}

extern "C" int ogv_mbconv_bn2_finalize(const float* stats, const float* gate, const float* dpool, float* dgamma2,
                                       float* dbeta2, int B, int HW, int Cm, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(stats && gate && dpool && dgamma2 && dbeta2 && Cm > 0 && HW > 0, "mbconv_bn2_finalize: bad args");
  mbconv_bn2_finalize_kernel<<<ogv_ceil_div(Cm, 32), dim3(32, 32), 0, (cudaStream_t)stream>>>(
      stats, gate, dpool, dgamma2, dbeta2, B, Cm, 1.f / (float)HW);
  return ogv_check_launch("mbconv_bn2_finalize");
}

extern "C" int ogv_bn_act_gate(const void* d_pre, const float* scale2, const float* shift2, const float* gate,
                               void* d_act, int B, int HW, int Cm, int act, int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(d_pre && scale2 && shift2 && gate && d_act && Cm % 8 == 0 && HW > 0, "bn_act_gate: bad args");
  const int nvb = ogv_ceil_div(Cm / 8, 32);
  dim3 grid(B, nvb, img_row_splits(B, nvb, HW));
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      bn_act_gate_kernel<T, ACT><<<grid, IMG_THREADS, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(d_pre), scale2, shift2, gate, reinterpret_cast<T*>(d_act), HW, Cm);
    });
    return ogv_check_launch("bn_act_gate");
  });
}

extern "C" int ogv_dw_bn2_bwd_apply(const void* dd_act, const void* d_pre, const float* gate, const float* dpool,
                                    const float* scale2, const float* shift2, const float* mean2, const float* rstd2,
                                    const float* gamma2, const float* dgamma2, const float* dbeta2, void* dd_pre,
                                    int B, int HW, int Cm, int act, int dtype, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(dd_act && d_pre && gate && dpool && scale2 && shift2 && mean2 && rstd2 && gamma2 && dgamma2 && dbeta2 &&
                  dd_pre,
              "dw_bn2_bwd_apply: null pointer");
  OGV_REQUIRE(Cm % 8 == 0 && HW > 0, "dw_bn2_bwd_apply: channels must be a multiple of 8");
  const int nvb = ogv_ceil_div(Cm / 8, 32);
  dim3 grid(B, nvb, img_row_splits(B, nvb, HW));
  OGV_DISPATCH_DTYPE(dtype, T, {
    OGV_DISPATCH_ACT(act, ACT, {
      dw_bn2_bwd_apply_kernel<T, ACT><<<grid, IMG_THREADS, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dd_act), reinterpret_cast<const T*>(d_pre), gate, dpool, scale2, shift2, mean2,
          rstd2, gamma2, dgamma2, dbeta2, reinterpret_cast<T*>(dd_pre), HW, Cm, 1.f / ((float)B * (float)HW));
    });
    return ogv_check_launch("dw_bn2_bwd_apply");
  });
}

extern "C" int ogv_se_mlp_supported(int Cm, int Cs, int dtype) {
  return dtype == OGV_BF16 && se_mlp_shape_ok(Cm, Cs) ? 1 : 0;
}

extern "C" int ogv_se_mlp_fwd(const float* pool, const void* w1t, const float* b1, const void* w2t, const float* b2,
                              void* pool_c, void* s1_pre, void* s1a, float* gate_pre, float* gate, int B, int Cm, int Cs,
                              int act, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(pool && w1t && b1 && w2t && b2 && pool_c && s1_pre && s1a && gate_pre && gate, "se_mlp_fwd: null pointer");
  OGV_REQUIRE(se_mlp_shape_ok(Cm, Cs), "se_mlp_fwd: unsupported shape Cm=%d Cs=%d", Cm, Cs);
  const size_t smem = se_mlp_smem(Cm, Cs);
  OGV_DISPATCH_ACT(act, ACT, {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(se_mlp_fwd_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    se_mlp_fwd_kernel<ACT><<<ogv_ceil_div(B, SE_RB), SE_THREADS, smem, (cudaStream_t)stream>>>(
        pool, reinterpret_cast<const bf16*>(w1t), b1, reinterpret_cast<const bf16*>(w2t), b2,
        reinterpret_cast<bf16*>(pool_c), reinterpret_cast<bf16*>(s1_pre), reinterpret_cast<bf16*>(s1a), gate_pre, gate, B,
        Cm, Cs);
  });
  return ogv_check_launch("se_mlp_fwd");
}

extern "C" int ogv_se_mlp_bwd(const float* dgate, const float* gate_pre, const void* s1_pre, const void* w2,
                              const void* w1, void* dgate_c, void* ds1_pre, float* dpool, int B, int Cm, int Cs, int act,
                              void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(dgate && gate_pre && s1_pre && w2 && w1 && dgate_c && ds1_pre && dpool, "se_mlp_bwd: null pointer");
  OGV_REQUIRE(se_mlp_shape_ok(Cm, Cs), "se_mlp_bwd: unsupported shape Cm=%d Cs=%d", Cm, Cs);
  const size_t smem = se_mlp_smem(Cm, Cs);
  OGV_DISPATCH_ACT(act, ACT, {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(se_mlp_bwd_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    se_mlp_bwd_kernel<ACT><<<ogv_ceil_div(B, SE_RB), SE_THREADS, smem, (cudaStream_t)stream>>>(
        dgate, gate_pre, reinterpret_cast<const bf16*>(s1_pre), reinterpret_cast<const bf16*>(w2),
        reinterpret_cast<const bf16*>(w1), reinterpret_cast<bf16*>(dgate_c), reinterpret_cast<bf16*>(ds1_pre), dpool, B,
        Cm, Cs);
  });
  return ogv_check_launch("se_mlp_bwd");
}
