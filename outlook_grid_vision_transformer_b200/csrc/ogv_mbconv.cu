// MBConv interior: depthwise 3x3 stencil with BatchNorm-affine + activation fused on load and batch
// statistics fused on store, SqueezeExcite pooling / gating, and their backward passes.
// Reference: mbc_conv.py:9-27 (SqueezeExcite), :44-98 (MBConv); math in SURVEY Appendix A.3/A.4.
//
// Depthwise kernels are persistent: a CTA owns one 32-channel chunk and walks spatial tiles
// (halo staged in shared memory, activation evaluated once per element), keeping the per-channel
// statistics / filter-gradient partial sums in registers until a single flush.
#include "ogv_common.cuh"
#include "ogv_reduce.cuh"
#include "../../include/ogv.h"

namespace {

constexpr int DW_CH = 32;        // channels per CTA chunk
constexpr int DW_PAD = 36;       // smem floats per position (32 + 4: conflict-free float4 access)
constexpr int DW_MAXPOS = 340;   // halo positions that fit 48 KB of static smem
constexpr int DW_THREADS = 256;
constexpr int DW_SMEM_BYTES = DW_MAXPOS * DW_PAD * 4;

template <typename K>
int dw_smem_optin(K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES);
  if (e != cudaSuccess) {
    ogv_set_error("dwconv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return OGV_ERR_CUDA;
  }
  return OGV_OK;
}

struct DwTiling {
  int TH, TW, NI;          // tile rows / cols / images per tile
  int tiles_h, tiles_w;    // tiles per image (1x1 when NI > 1)
  long long ntiles;
};

DwTiling dw_tiling(int B, int H, int W) {
  DwTiling t;
  t.TW = W < 32 ? W : 32;
  int max_th = DW_MAXPOS / (t.TW + 2) - 2;
  t.TH = H < max_th ? H : max_th;
  if (t.TH >= 4 && t.TH < H) t.TH &= ~3;
  t.tiles_h = (H + t.TH - 1) / t.TH;
  t.tiles_w = (W + t.TW - 1) / t.TW;
  t.NI = 1;
  if (t.tiles_h == 1 && t.tiles_w == 1) {
    t.NI = DW_MAXPOS / ((t.TH + 2) * (t.TW + 2));
    if (t.NI < 1) t.NI = 1;
    if (t.NI > B) t.NI = B > 0 ? B : 1;
  }
  long long groups = (B + t.NI - 1) / t.NI;
  t.ntiles = groups * t.tiles_h * t.tiles_w;
  return t;
}

struct DwGeom {
  int B, H, W, Cm;
  int TH, TW, NI, tiles_h, tiles_w;
  long long ntiles;
  int nchunks, nworkers;
};

// Fill the halo tile: smem[pos][ch] = f(global) for in-image positions, 0 elsewhere.
// APPLY: apply scale/shift + activation (forward input) ; otherwise raw copy (backward gradient).
template <typename T, bool APPLY>
__device__ __forceinline__ void dw_fill_tile(float* __restrict__ tile, const T* __restrict__ src,
                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                             const DwGeom& g, int b0, int h0, int w0, int c0, int act) {
  const int ncv8 = DW_CH / 8;
  const int cv = threadIdx.x % ncv8;
  const int c = c0 + cv * 8;
  const bool cvalid = c < g.Cm;
  float sc[8], sh[8];
  if (APPLY && cvalid) {
    ld8(scale + c, sc);
    ld8(shift + c, sh);
  }
  const int tw2 = g.TW + 2, th2 = g.TH + 2;
  const int npos = g.NI * th2 * tw2;
  for (int i = threadIdx.x; i < npos * ncv8; i += DW_THREADS) {
    const int pos = i / ncv8;
    const int img = pos / (th2 * tw2);
    const int rem = pos % (th2 * tw2);
    const int gh = h0 - 1 + rem / tw2, gw = w0 - 1 + rem % tw2;
    const int b = b0 + img;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    if (cvalid && b < g.B && gh >= 0 && gh < g.H && gw >= 0 && gw < g.W) {
      ld8(src + (((long long)b * g.H + gh) * g.W + gw) * g.Cm + c, v);
      if (APPLY) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = act_apply(act, v[k] * sc[k] + sh[k]);
      }
    }
    float4* dst = reinterpret_cast<float4*>(tile + pos * DW_PAD + cv * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__device__ __forceinline__ void dw_decode_tile(const DwGeom& g, long long t, int& b0, int& h0, int& w0) {
  const int tw_i = (int)(t % g.tiles_w);
  const int th_i = (int)((t / g.tiles_w) % g.tiles_h);
  const long long grp = t / ((long long)g.tiles_w * g.tiles_h);
  b0 = (int)(grp * g.NI);
  h0 = th_i * g.TH;
  w0 = tw_i * g.TW;
}

// ------------------------------------------------------------------------------------------------
// forward: d_pre[p,c] = sum_t w[c,t] * act(scale1*e_pre + shift1)[p + d_t, c]; stats of d_pre
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(DW_THREADS, 2) dwconv_fwd_kernel(const T* __restrict__ e_pre,
                                                                const float* __restrict__ scale1,
                                                                const float* __restrict__ shift1,
                                                                const float* __restrict__ wgt, T* __restrict__ d_pre,
                                                                float* __restrict__ sum2, float* __restrict__ sumsq2,
                                                                DwGeom g, int act) {
  extern __shared__ __align__(16) float tile[];  // DW_MAXPOS * DW_PAD floats
  __shared__ float s_sum[DW_CH], s_sq[DW_CH];
  constexpr int R = 4;
  const int chunk = blockIdx.x % g.nchunks;
  const int worker = blockIdx.x / g.nchunks;
  const int c0 = chunk * DW_CH;
  const int cv = threadIdx.x % 8;  // 4-channel group inside the chunk
  const int c = c0 + cv * 4;
  const bool cvalid = c < g.Cm;
  if (threadIdx.x < DW_CH) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }

  float w[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) w[t][k] = cvalid ? wgt[(c + k) * 9 + t] : 0.f;
  float st_s[4] = {0.f, 0.f, 0.f, 0.f}, st_q[4] = {0.f, 0.f, 0.f, 0.f};

  const int tw2 = g.TW + 2, th2 = g.TH + 2;
  const int nstrips = (g.TH + R - 1) / R;
  const int nitems = g.NI * nstrips * g.TW * 8;

  for (long long t = worker; t < g.ntiles; t += g.nworkers) {
    int b0, h0, w0;
    dw_decode_tile(g, t, b0, h0, w0);
    __syncthreads();
    dw_fill_tile<T, true>(tile, e_pre, scale1, shift1, g, b0, h0, w0, c0, act);
    __syncthreads();
    for (int it = threadIdx.x; it < nitems; it += DW_THREADS) {
      // it = ((img * nstrips + strip) * TW + x) * 8 + cv
      const int x = (it / 8) % g.TW;
      const int strip = (it / (8 * g.TW)) % nstrips;
      const int img = it / (8 * g.TW * nstrips);
      const int r0 = strip * R;
      const int b = b0 + img;
      if (!cvalid || b >= g.B || w0 + x >= g.W) continue;
      float acc[R][4];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[r][k] = 0.f;
      const float* base = tile + ((img * th2) * tw2) * DW_PAD + cv * 4;
#pragma unroll
      for (int ry = 0; ry < R + 2; ++ry) {  // input row (tile coords incl. halo) r0 + ry
        const int trow = r0 + ry;
        if (trow >= th2) break;
        float in[3][4];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          float4 v = *reinterpret_cast<const float4*>(base + (trow * tw2 + x + dx) * DW_PAD);
          in[dx][0] = v.x; in[dx][1] = v.y; in[dx][2] = v.z; in[dx][3] = v.w;
        }
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          const int r = ry - ki;  // output row inside the strip that sees this input row through tap row ki
          if (r >= 0 && r < R) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int k = 0; k < 4; ++k) acc[r][k] = fmaf(w[ki * 3 + dx][k], in[dx][k], acc[r][k]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int gh = h0 + r0 + r;
        if (r0 + r < g.TH && gh < g.H) {
          T* dst = d_pre + (((long long)b * g.H + gh) * g.W + (w0 + x)) * g.Cm + c;
          stv<4>(dst, acc[r]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float s = round_to<T>(acc[r][k]);  // statistics of the value as stored (rounded)
            st_s[k] += s;
            st_q[k] += s * s;
          }
        }
      }
    }
  }
  if (cvalid) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&s_sum[cv * 4 + k], st_s[k]);
      atomicAdd(&s_sq[cv * 4 + k], st_q[k]);
    }
  }
  __syncthreads();
  if (threadIdx.x < DW_CH && c0 + threadIdx.x < g.Cm) {
    if (sum2) atomicAdd(sum2 + c0 + threadIdx.x, s_sum[threadIdx.x]);
    if (sumsq2) atomicAdd(sumsq2 + c0 + threadIdx.x, s_sq[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: de_act[p] = sum_t w[t] * G[p - d_t];  du1 = de_act * act'(u1[p]);
//           dw[c,t] += sum_p e_act[p] * G[p - d_t];  dbeta1 += du1; dgamma1 += du1 * xhat1
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(DW_THREADS, 2) dwconv_bwd_kernel(
    const T* __restrict__ dd_pre, const T* __restrict__ e_pre, const float* __restrict__ scale1,
    const float* __restrict__ shift1, const float* __restrict__ mean1, const float* __restrict__ rstd1,
    const float* __restrict__ wgt, T* __restrict__ du1, float* __restrict__ dwgt, float* __restrict__ dgamma1,
    float* __restrict__ dbeta1, DwGeom g, int act) {
  extern __shared__ __align__(16) float tile[];  // DW_MAXPOS * DW_PAD floats
  __shared__ float s_dw[DW_CH * 9], s_db[DW_CH], s_dg[DW_CH];
  constexpr int R = 2;
  const int chunk = blockIdx.x % g.nchunks;
  const int worker = blockIdx.x / g.nchunks;
  const int c0 = chunk * DW_CH;
  const int cv = threadIdx.x % 8;
  const int c = c0 + cv * 4;
  const bool cvalid = c < g.Cm;
  for (int i = threadIdx.x; i < DW_CH * 9; i += DW_THREADS) s_dw[i] = 0.f;
  if (threadIdx.x < DW_CH) { s_db[threadIdx.x] = 0.f; s_dg[threadIdx.x] = 0.f; }

  float w[9][4], dwa[9][4], sc[4], sh[4], mu[4], rs[4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      w[t][k] = cvalid ? wgt[(c + k) * 9 + t] : 0.f;
      dwa[t][k] = 0.f;
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sc[k] = cvalid ? scale1[c + k] : 0.f;
    sh[k] = cvalid ? shift1[c + k] : 0.f;
    mu[k] = cvalid ? mean1[c + k] : 0.f;
    rs[k] = cvalid ? rstd1[c + k] : 0.f;
  }
  float a_db[4] = {0.f, 0.f, 0.f, 0.f}, a_dg[4] = {0.f, 0.f, 0.f, 0.f};

  const int tw2 = g.TW + 2, th2 = g.TH + 2;
  const int nstrips = (g.TH + R - 1) / R;
  const int nitems = g.NI * nstrips * g.TW * 8;

  for (long long t = worker; t < g.ntiles; t += g.nworkers) {
    int b0, h0, w0;
    dw_decode_tile(g, t, b0, h0, w0);
    __syncthreads();
    dw_fill_tile<T, false>(tile, dd_pre, nullptr, nullptr, g, b0, h0, w0, c0, act);
    __syncthreads();
    for (int it = threadIdx.x; it < nitems; it += DW_THREADS) {
      const int x = (it / 8) % g.TW;
      const int strip = (it / (8 * g.TW)) % nstrips;
      const int img = it / (8 * g.TW * nstrips);
      const int r0 = strip * R;
      const int b = b0 + img;
      if (!cvalid || b >= g.B || w0 + x >= g.W) continue;
      float de[R][4], ea[R][4], da[R][4], xh[R][4];
      bool rvalid[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int gh = h0 + r0 + r;
        rvalid[r] = (r0 + r < g.TH) && (gh < g.H);
#pragma unroll
        for (int k = 0; k < 4; ++k) { de[r][k] = 0.f; ea[r][k] = 0.f; da[r][k] = 0.f; xh[r][k] = 0.f; }
        if (rvalid[r]) {
          float ev[4];
          ldv<4>(e_pre + (((long long)b * g.H + gh) * g.W + (w0 + x)) * g.Cm + c, ev);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float u = ev[k] * sc[k] + sh[k];
            ea[r][k] = act_apply(act, u);
            da[r][k] = act_grad(act, u);
            xh[r][k] = (ev[k] - mu[k]) * rs[k];
          }
        }
      }
      const float* base = tile + ((img * th2) * tw2) * DW_PAD + cv * 4;
#pragma unroll
      for (int ry = 0; ry < R + 2; ++ry) {  // gradient row (tile coords incl. halo) r0 + ry  <->  image row h0+r0+ry-1
        const int trow = r0 + ry;
        if (trow >= th2) break;
        float gr[3][4];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          float4 v = *reinterpret_cast<const float4*>(base + (trow * tw2 + x + dx) * DW_PAD);
          gr[dx][0] = v.x; gr[dx][1] = v.y; gr[dx][2] = v.z; gr[dx][3] = v.w;
        }
        // output row r (strip coords) sits at tile row r0 + r + 1; gradient row offset = (ry - 1) - r = -(ki - 1)
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          const int r = ry - 1 + (ki - 1);
          if (r >= 0 && r < R) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              // gradient column offset = dx - 1 = -(kj - 1)  ->  kj = 2 - dx
              const int tt = ki * 3 + (2 - dx);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                de[r][k] = fmaf(w[tt][k], gr[dx][k], de[r][k]);
                dwa[tt][k] = fmaf(ea[r][k], gr[dx][k], dwa[tt][k]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (rvalid[r]) {
          const int gh = h0 + r0 + r;
          float o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            o[k] = de[r][k] * da[r][k];
            a_db[k] += o[k];
            a_dg[k] += o[k] * xh[r][k];
          }
          stv<4>(du1 + (((long long)b * g.H + gh) * g.W + (w0 + x)) * g.Cm + c, o);
        }
      }
    }
  }
  if (cvalid) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&s_db[cv * 4 + k], a_db[k]);
      atomicAdd(&s_dg[cv * 4 + k], a_dg[k]);
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&s_dw[(cv * 4 + k) * 9 + t], dwa[t][k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < DW_CH * 9; i += DW_THREADS)
    if (c0 + i / 9 < g.Cm) atomicAdd(dwgt + (long long)c0 * 9 + i, s_dw[i]);
  if (threadIdx.x < DW_CH && c0 + threadIdx.x < g.Cm) {
    atomicAdd(dbeta1 + c0 + threadIdx.x, s_db[threadIdx.x]);
    atomicAdd(dgamma1 + c0 + threadIdx.x, s_dg[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------------
// SE squeeze and gate kernels (per image reductions)
// ------------------------------------------------------------------------------------------------
// MODE 0: out[b,c] = (1/HW) sum_p act(sc*x+sh)      MODE 1: out[b,c] = sum_p y[p,c] * act(sc*x+sh)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) se_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift, float* __restrict__ out,
                                                        int HW, int Cm, int act) {
  __shared__ float red[8][32][9];
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
    for (int p = ry; p < HW; p += 8) {
      const long long off = ((long long)b * HW + p) * Cm + cv * 8;
      float v[8];
      ld8(x + off, v);
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += act_apply(act, v[k] * sc[k] + sh[k]);
      } else {
        float yv[8];
        ld8(y + off, yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += yv[k] * act_apply(act, v[k] * sc[k] + sh[k]);
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
      for (int j = 0; j < 8; ++j) s += red[j][cx][k];
      out[(long long)b * Cm + cv * 8 + k] = s * norm;
    }
  }
}

template <typename T>
__global__ void bn_act_gate_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                   const float* __restrict__ shift, const float* __restrict__ gate,
                                   T* __restrict__ out, long long nvec, int nv, int HW, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % nv);
    const long long row = i / nv;
    const long long b = row / HW;
    float v[8], sc[8], sh[8], gt[8];
    ld8(x + i * 8, v);
    ld8(scale + cv * 8, sc);
    ld8(shift + cv * 8, sh);
    ld8(gate + (b * nv + cv) * 8, gt);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = act_apply(act, v[k] * sc[k] + sh[k]) * gt[k];
    st8(out + i * 8, v);
  }
}

// du = (dd_act*gate + dpool/HW) * act'(u), u = sc*x + sh ; xhat = (x - mean)*rstd
template <typename T>
__device__ __forceinline__ void dw_bn2_du(const T* dd_act, const T* d_pre, const float* gate, const float* dpool,
                                          const float (&sc)[8], const float (&sh)[8], const float (&mu)[8],
                                          const float (&rs)[8], long long vec_idx, long long b, int nv, int cv,
                                          float inv_hw, int act, float (&du)[8], float (&xh)[8]) {
  float g[8], x[8], gt[8], dp[8];
  ld8(dd_act + vec_idx * 8, g);
  ld8(d_pre + vec_idx * 8, x);
  ld8(gate + (b * nv + cv) * 8, gt);
  ld8(dpool + (b * nv + cv) * 8, dp);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float u = x[k] * sc[k] + sh[k];
    du[k] = (g[k] * gt[k] + dp[k] * inv_hw) * act_grad(act, u);
    xh[k] = (x[k] - mu[k]) * rs[k];
  }
}

template <typename T>
__global__ void __launch_bounds__(COLREDUCE_THREADS) dw_bn2_bwd_reduce_kernel(const T* __restrict__ dd_act, const T* __restrict__ d_pre,
                                         const float* __restrict__ gate, const float* __restrict__ dpool,
                                         const float* __restrict__ scale2, const float* __restrict__ shift2,
                                         const float* __restrict__ mean2, const float* __restrict__ rstd2,
                                         float* __restrict__ dgamma2, float* __restrict__ dbeta2, long long M, int nv,
                                         int HW, int act) {
  float acc[2][8];
  colreduce_init(acc);
  float sc[8], sh[8], mu[8], rs[8];
  {
    const int cv0 = threadIdx.x % nv;
    ld8(scale2 + cv0 * 8, sc);
    ld8(shift2 + cv0 * 8, sh);
    ld8(mean2 + cv0 * 8, mu);
    ld8(rstd2 + cv0 * 8, rs);
  }
  const float inv_hw = 1.f / (float)HW;
  COLREDUCE_LOOP(M, nv, row, cv) {
    float du[8], xh[8];
    dw_bn2_du(dd_act, d_pre, gate, dpool, sc, sh, mu, rs, row * nv + cv, row / HW, nv, cv, inv_hw, act, du, xh);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[0][k] += du[k];
      acc[1][k] += du[k] * xh[k];
    }
  }
  float* outs[2] = {dbeta2, dgamma2};
  colreduce_finish<2>(acc, outs, nv);
}

template <typename T>
__global__ void dw_bn2_bwd_apply_kernel(const T* __restrict__ dd_act, const T* __restrict__ d_pre,
                                        const float* __restrict__ gate, const float* __restrict__ dpool,
                                        const float* __restrict__ scale2, const float* __restrict__ shift2,
                                        const float* __restrict__ mean2, const float* __restrict__ rstd2,
                                        const float* __restrict__ gamma2, const float* __restrict__ dgamma2,
                                        const float* __restrict__ dbeta2, T* __restrict__ dd_pre, long long nvec,
                                        int nv, int HW, float inv_n, int act) {
  const float inv_hw = 1.f / (float)HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % nv);
    const long long row = i / nv;
    float sc[8], sh[8], mu[8], rs[8], gm[8], dg[8], db[8], du[8], xh[8], o[8];
    ld8(scale2 + cv * 8, sc);
    ld8(shift2 + cv * 8, sh);
    ld8(mean2 + cv * 8, mu);
    ld8(rstd2 + cv * 8, rs);
    ld8(gamma2 + cv * 8, gm);
    ld8(dgamma2 + cv * 8, dg);
    ld8(dbeta2 + cv * 8, db);
    dw_bn2_du(dd_act, d_pre, gate, dpool, sc, sh, mu, rs, i, row / HW, nv, cv, inv_hw, act, du, xh);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = gm[k] * rs[k] * (du[k] - db[k] * inv_n - xh[k] * dg[k] * inv_n);
    st8(dd_pre + i * 8, o);
  }
}

inline int flat_grid(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)ogv_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int make_geom(int B, int H, int W, int Cm, DwGeom* g) {
  DwTiling t = dw_tiling(B, H, W);
  g->B = B; g->H = H; g->W = W; g->Cm = Cm;
  g->TH = t.TH; g->TW = t.TW; g->NI = t.NI; g->tiles_h = t.tiles_h; g->tiles_w = t.tiles_w; g->ntiles = t.ntiles;
  g->nchunks = (Cm + DW_CH - 1) / DW_CH;
  long long want = ((long long)ogv_num_sms() * 4 + g->nchunks - 1) / g->nchunks;
  if (want > t.ntiles) want = t.ntiles;
  if (want < 1) want = 1;
  g->nworkers = (int)want;
  if (t.TH < 1 || t.TW < 1 || (t.TH + 2) * (t.TW + 2) * t.NI > DW_MAXPOS) return -1;
  return 0;
}

}  // namespace

extern "C" int ogv_dwconv_fwd(const void* e_pre, const float* scale1, const float* shift1, const float* w,
                              void* d_pre, float* sum2, float* sumsq2, int B, int H, int W, int Cm, int act,
                              int dtype, void* stream) {
  OGV_REQUIRE(e_pre && scale1 && shift1 && w && d_pre, "dwconv_fwd: null pointer");
  OGV_REQUIRE(Cm > 0 && Cm % 8 == 0 && H > 0 && W > 0, "dwconv_fwd: channels must be a multiple of 8");
  if (B == 0) return OGV_OK;
  DwGeom g;
  if (make_geom(B, H, W, Cm, &g)) { ogv_set_error("dwconv_fwd: cannot tile %dx%d", H, W); return OGV_ERR_UNSUPPORTED; }
  OGV_DISPATCH_DTYPE(dtype, T, {
    if (int rc = dw_smem_optin(dwconv_fwd_kernel<T>)) return rc;
    dwconv_fwd_kernel<T><<<g.nchunks * g.nworkers, DW_THREADS, DW_SMEM_BYTES, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(e_pre), scale1, shift1, w, reinterpret_cast<T*>(d_pre), sum2, sumsq2, g, act);
    return ogv_check_launch("dwconv_fwd");
  });
}

extern "C" int ogv_dwconv_bwd(const void* dd_pre, const void* e_pre, const float* scale1, const float* shift1,
                              const float* mean1, const float* rstd1, const float* w, void* du1, float* dw,
                              float* dgamma1, float* dbeta1, int B, int H, int W, int Cm, int act, int dtype,
                              void* stream) {
  OGV_REQUIRE(dd_pre && e_pre && scale1 && shift1 && mean1 && rstd1 && w && du1 && dw && dgamma1 && dbeta1,
              "dwconv_bwd: null pointer");
  OGV_REQUIRE(Cm > 0 && Cm % 8 == 0 && H > 0 && W > 0, "dwconv_bwd: channels must be a multiple of 8");
  if (B == 0) return OGV_OK;
  DwGeom g;
  if (make_geom(B, H, W, Cm, &g)) { ogv_set_error("dwconv_bwd: cannot tile %dx%d", H, W); return OGV_ERR_UNSUPPORTED; }
  OGV_DISPATCH_DTYPE(dtype, T, {
    if (int rc = dw_smem_optin(dwconv_bwd_kernel<T>)) return rc;
    dwconv_bwd_kernel<T><<<g.nchunks * g.nworkers, DW_THREADS, DW_SMEM_BYTES, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(dd_pre), reinterpret_cast<const T*>(e_pre), scale1, shift1, mean1, rstd1, w,
        reinterpret_cast<T*>(du1), dw, dgamma1, dbeta1, g, act);
    return ogv_check_launch("dwconv_bwd");
  });
}

extern "C" int ogv_se_pool(const void* d_pre, const float* scale2, const float* shift2, float* pool, int B, int HW,
                           int Cm, int act, int dtype, void* stream) {
  OGV_REQUIRE(d_pre && scale2 && shift2 && pool && Cm % 8 == 0 && HW > 0, "se_pool: bad args");
  if (B == 0) return OGV_OK;
  dim3 grid(B, ogv_ceil_div(Cm / 8, 32));
  OGV_DISPATCH_DTYPE(dtype, T, {
    se_reduce_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(d_pre), nullptr, scale2,
                                                                   shift2, pool, HW, Cm, act);
    return ogv_check_launch("se_pool");
  });
}

extern "C" int ogv_se_bwd_reduce(const void* dd_act, const void* d_pre, const float* scale2, const float* shift2,
                                 float* dgate, int B, int HW, int Cm, int act, int dtype, void* stream) {
  OGV_REQUIRE(dd_act && d_pre && scale2 && shift2 && dgate && Cm % 8 == 0 && HW > 0, "se_bwd_reduce: bad args");
  if (B == 0) return OGV_OK;
  dim3 grid(B, ogv_ceil_div(Cm / 8, 32));
  OGV_DISPATCH_DTYPE(dtype, T, {
    se_reduce_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(d_pre),
                                                                   reinterpret_cast<const T*>(dd_act), scale2, shift2,
                                                                   dgate, HW, Cm, act);
    return ogv_check_launch("se_bwd_reduce");
  });
}

extern "C" int ogv_bn_act_gate(const void* d_pre, const float* scale2, const float* shift2, const float* gate,
                               void* d_act, int B, int HW, int Cm, int act, int dtype, void* stream) {
  OGV_REQUIRE(d_pre && scale2 && shift2 && gate && d_act && Cm % 8 == 0 && HW > 0, "bn_act_gate: bad args");
  long long nvec = (long long)B * HW * (Cm / 8);
  if (nvec == 0) return OGV_OK;
  OGV_DISPATCH_DTYPE(dtype, T, {
    bn_act_gate_kernel<T><<<flat_grid(nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(d_pre), scale2, shift2, gate, reinterpret_cast<T*>(d_act), nvec, Cm / 8, HW, act);
    return ogv_check_launch("bn_act_gate");
  });
}

extern "C" int ogv_dw_bn2_bwd(int pass, const void* dd_act, const void* d_pre, const float* gate,
                              const float* dpool, const float* scale2, const float* shift2, const float* mean2,
                              const float* rstd2, const float* gamma2, float* dgamma2, float* dbeta2, void* dd_pre,
                              int B, int HW, int Cm, int act, int dtype, void* stream) {
  OGV_REQUIRE(dd_act && d_pre && gate && dpool && scale2 && shift2 && mean2 && rstd2 && dgamma2 && dbeta2,
              "dw_bn2_bwd: null pointer");
  OGV_REQUIRE(Cm % 8 == 0 && HW > 0, "dw_bn2_bwd: channels must be a multiple of 8");
  const long long M = (long long)B * HW;
  if (M == 0) return OGV_OK;
  const int nv = Cm / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (pass == 0) {
    ColReduceCfg cfg;
    if (!colreduce_config(M, nv, &cfg)) { ogv_set_error("dw_bn2_bwd: Cm=%d too wide", Cm); return OGV_ERR_UNSUPPORTED; }
    OGV_DISPATCH_DTYPE(dtype, T, {
      dw_bn2_bwd_reduce_kernel<T><<<cfg.grid, cfg.block, 0, st>>>(
          reinterpret_cast<const T*>(dd_act), reinterpret_cast<const T*>(d_pre), gate, dpool, scale2, shift2, mean2,
          rstd2, dgamma2, dbeta2, M, nv, HW, act);
      return ogv_check_launch("dw_bn2_bwd_reduce");
    });
  }
  OGV_REQUIRE(gamma2 && dd_pre, "dw_bn2_bwd: apply pass needs gamma2 and dd_pre");
  OGV_DISPATCH_DTYPE(dtype, T, {
    dw_bn2_bwd_apply_kernel<T><<<flat_grid(M * nv, 256), 256, 0, st>>>(
        reinterpret_cast<const T*>(dd_act), reinterpret_cast<const T*>(d_pre), gate, dpool, scale2, shift2, mean2,
        rstd2, gamma2, dgamma2, dbeta2, reinterpret_cast<T*>(dd_pre), M * nv, nv, HW, 1.f / (float)M, act);
    return ogv_check_launch("dw_bn2_bwd_apply");
  });
}
