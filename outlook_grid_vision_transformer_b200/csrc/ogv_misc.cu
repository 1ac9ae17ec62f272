// Library plumbing (error text, launch checks, GEMM engine selection) and the small layout /
// elementwise kernels around the hot path.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "ogv_gemm.cuh"
#include "ogv_reduce.cuh"

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void ogv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static int g_sync_mode = -1;
int ogv_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ogv_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return OGV_ERR_CUDA;
  }
  if (g_sync_mode < 0) {
    const char* s = getenv("OGV_SYNC");
    g_sync_mode = (s && s[0] == '1') ? 1 : 0;
  }
  if (g_sync_mode == 1) {  // debugging aid only: surfaces asynchronous faults at the faulting launch
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      ogv_set_error("%s: execution failed: %s", what, cudaGetErrorString(e));
      return OGV_ERR_CUDA;
    }
  }
  return OGV_OK;
}
int ogv_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

extern "C" int ogv_version(void) { return 100; }
extern "C" const char* ogv_last_error(void) { return g_err; }
extern "C" int ogv_sm_count(void) { return ogv_num_sms(); }

extern "C" int ogv_gemm(const ogv_gemm_args* args, int engine, void* stream) {
  if (!args) { ogv_set_error("ogv_gemm: null args"); return OGV_ERR_ARG; }
  const ogv_gemm_args& a = *args;
  OGV_REQUIRE(a.M >= 0 && a.N >= 0 && a.K >= 0, "ogv_gemm: negative extent");
  if (a.M == 0 || a.N == 0) return OGV_OK;  // empty problem: nothing to do (operands may legitimately be null)
  OGV_REQUIRE(a.A && a.B && a.D, "ogv_gemm: null operand");
  OGV_REQUIRE(!(a.split_k > 1 && !a.accumulate), "ogv_gemm: split_k>1 requires accumulate");
  OGV_REQUIRE(!(a.accumulate && a.out_dtype != OGV_F32), "ogv_gemm: accumulate requires fp32 output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (engine == OGV_ENGINE_AUTO) {
    static int force_simt = -1;
    if (force_simt < 0) {
      const char* s = getenv("OGV_FORCE_SIMT");
      force_simt = (s && s[0] == '1') ? 1 : 0;
    }
    engine = (!force_simt && ogv_gemm_tc_supported(a, nullptr)) ? OGV_ENGINE_TC : OGV_ENGINE_SIMT;
  }
  if (engine == OGV_ENGINE_TC) return ogv_gemm_tc(a, st);
  if (engine == OGV_ENGINE_SIMT) {
    if (a.row_sum) {  // the FFMA engine has no row-sum path: one column-sum pass over the operand instead
      OGV_REQUIRE(a.a_rs == 1, "ogv_gemm: row_sum on the FFMA engine needs the weight-gradient operand layout (a_rs == 1)");
      if (int rc = ogv_colsum(a.A, a.a_cs, a.row_sum, a.K, a.M, a.in_dtype, stream)) return rc;
      ogv_gemm_args b = a;
      b.row_sum = nullptr;
      return ogv_gemm_simt(b, st);
    }
    return ogv_gemm_simt(a, st);
  }
  ogv_set_error("ogv_gemm: unknown engine %d", engine);
  return OGV_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------
// NCHW <-> NHWC per image: [C, HW] <-> [HW, C]  (32x32 smem tile transpose)
// ---------------------------------------------------------------------------------------------
namespace {

template <typename T>
__global__ void transpose_batched_kernel(const T* __restrict__ src, T* __restrict__ dst, int R, int Ccols) {
  // src: [B][R][Ccols] -> dst: [B][Ccols][R]
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * R * Ccols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Ccols) tile[i][threadIdx.x] = ld1(src + base + (long long)r * Ccols + c);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Ccols) st1(dst + base + (long long)c * R + r, tile[threadIdx.x][i]);
  }
}

template <typename T>
int transpose_batched(const void* src, void* dst, int B, int R, int Ccols, cudaStream_t st) {
  if (B == 0 || R == 0 || Ccols == 0) return OGV_OK;
  dim3 grid(ogv_ceil_div(Ccols, 32), ogv_ceil_div(R, 32), B);
  if (grid.y > 65535 || grid.z > 65535) {
    ogv_set_error("transpose: extent too large (R=%d B=%d)", R, B);
    return OGV_ERR_UNSUPPORTED;
  }
  transpose_batched_kernel<T><<<grid, dim3(32, 8), 0, st>>>(reinterpret_cast<const T*>(src),
                                                            reinterpret_cast<T*>(dst), R, Ccols);
  return ogv_check_launch("transpose");
}

template <typename T>
__global__ void cast_transpose_kernel(const float* __restrict__ src, T* __restrict__ dst, long long ld_dst,
                                      T* __restrict__ dst_t, long long ld_dst_t, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) {
      float v = src[(long long)r * cols + c];
      tile[i][threadIdx.x] = v;
      if (dst) st1(dst + (long long)r * ld_dst + c, v);
    }
  }
  __syncthreads();
  if (dst_t) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      int c = c0 + i, r = r0 + threadIdx.x;
      if (r < rows && c < cols) st1(dst_t + (long long)c * ld_dst_t + r, tile[threadIdx.x][i]);
    }
  }
}

// every weight of a model in one launch: CTA t handles 32x32 tile t of the batch; the owning item is found by a
// binary search over the items' first-tile indices (a few hundred items at most).
template <typename T>
__device__ __forceinline__ void cast_tile(const ogv_cast_item& it, int tx, int ty, float (&tile)[32][33]) {
  T* dst = reinterpret_cast<T*>(it.dst);
  T* dst_t = reinterpret_cast<T*>(it.dst_t);
  const int c0 = tx * 32, r0 = ty * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < it.rows && c < it.cols) {
      const float v = it.src[(long long)r * it.cols + c];
      tile[i][threadIdx.x] = v;
      if (dst) st1(dst + (long long)r * it.ld_dst + c, v);
    }
  }
  __syncthreads();
  if (dst_t) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < it.rows && c < it.cols) st1(dst_t + (long long)c * it.ld_dst_t + r, tile[threadIdx.x][i]);
    }
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256) cast_batch_kernel(const ogv_cast_item* __restrict__ items, int n_items,
                                                         int total_tiles) {
  __shared__ float tile[32][33];
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].tile0 <= t) lo = mid; else hi = mid - 1;
    }
    const ogv_cast_item it = items[lo];
    const int tiles_x = (it.cols + 31) >> 5;
    const int lt = t - it.tile0;
    const int ty = lt / tiles_x, tx = lt - ty * tiles_x;
    if (it.dst_dtype == OGV_BF16) cast_tile<bf16>(it, tx, ty, tile);
    else cast_tile<float>(it, tx, ty, tile);
  }
}

// fp32 -> three bf16 planes for the "bf16 x 3" tensor-core emulation of an fp32 product:
//   a = hi + lo (+ 2^-17 a),  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi   (every bf16 product is exact in fp32)
// pattern 0 writes (hi, hi, lo), pattern 1 writes (hi, lo, hi), so that the element-wise product of a pattern-0
// and a pattern-1 operand along a 3x longer reduction axis is exactly that sum.  plane_stride = distance between
// planes in dst elements (cols: planes side by side in a [rows, 3*cols] matrix; rows*ld_dst: stacked).
__global__ void split3_kernel(const float* __restrict__ src, long long ld_src, bf16* __restrict__ dst, long long ld_dst,
                              long long plane_stride, long long rows, int cols, int pattern) {
  // one thread per 8 consecutive columns (cols % 8 == 0): two 16-byte loads, three 16-byte stores
  const unsigned vpr = (unsigned)cols >> 3;
  const long long nvec = rows * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const unsigned r = (unsigned)i / vpr;  // host guarantees rows * cols / 8 < 2^32
    const unsigned c = ((unsigned)i - r * vpr) << 3;
    float a[8];
    ld8(src + (long long)r * ld_src + c, a);
    float hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      hi[k] = round_to<bf16>(a[k]);
      lo[k] = a[k] - hi[k];
    }
    bf16* o = dst + (long long)r * ld_dst + c;
    st8(o, hi);
    st8(o + plane_stride, pattern == 0 ? hi : lo);
    st8(o + 2 * plane_stride, pattern == 0 ? lo : hi);
  }
}

template <typename T>
__global__ void rowscale_kernel(const T* __restrict__ x, const float* __restrict__ scale, T* __restrict__ y,
                                long long nvec, int vec_per_row, int rows_per_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    long long row = i / vec_per_row;
    float s = scale[row / rows_per_scale];
    float v[8];
    ld8(x + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= s;
    st8(y + i * 8, v);
  }
}

template <typename T>
__global__ void mul_dact_kernel(const T* __restrict__ a, const T* __restrict__ pre, T* __restrict__ out, long long n,
                                int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st1(out + i, ld1(a + i) * act_grad(act, ld1(pre + i)));
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st1(y + i, ld1(a + i) + ld1(b + i));
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i] * (1.f - lr * wd);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * mi / denom;
  }
}

struct ColSumF {
  template <typename T>
  __device__ __forceinline__ void operator()(const T* x, long long ld, long long row, int col, float (&acc)[1][8]) const {
    float v[8];
    ld8(x + row * ld + col, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[0][i] += v[i];
  }
};

template <typename T>
__global__ void __launch_bounds__(COLREDUCE_THREADS) colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out, long long M, int nv) {
  float acc[1][8];
  colreduce_init(acc);
  COLREDUCE_LOOP(M, nv, row, cv) {
    float v[8];
    ld8(x + row * ld + cv * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[0][i] += v[i];
  }
  float* outs[1] = {out};
  colreduce_finish<1>(acc, outs, nv);
}

// y = x * scale[row / rows_per_scale] and out[n] += sum_m y[m,n] in one pass (DropPath backward feeding
// the bias gradient of the branch's last linear layer).
template <typename T>
__global__ void __launch_bounds__(COLREDUCE_THREADS, 3) rowscale_colsum_kernel(const T* __restrict__ x,
                                                                            const float* __restrict__ scale,
                                                                            T* __restrict__ y, float* __restrict__ out,
                                                                            long long M, int nv, unsigned rows_per_scale) {
  float acc[1][8];
  colreduce_init(acc);
  COLREDUCE_LOOP(M, nv, row, cv) {
    float v[8];
    ld8(x + (row * nv + cv) * 8, v);
    const float s = scale[(unsigned)row / rows_per_scale];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] *= s;
      acc[0][i] += v[i];
    }
    if (y) st8(y + (row * nv + cv) * 8, v);
  }
  float* outs[1] = {out};
  colreduce_finish<1>(acc, outs, nv);
}

template <typename T>
__global__ void __launch_bounds__(COLREDUCE_THREADS) colstats_kernel(const T* __restrict__ x, long long ld, float* __restrict__ sum,
                                float* __restrict__ sumsq, long long M, int nv) {
  float acc[2][8];
  colreduce_init(acc);
  COLREDUCE_LOOP(M, nv, row, cv) {
    float v[8];
    ld8(x + row * ld + cv * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[0][i] += v[i];
      acc[1][i] += v[i] * v[i];
    }
  }
  float* outs[2] = {sum, sumsq};
  colreduce_finish<2>(acc, outs, nv);
}

// ------------------------------------------------------------------------------------------------
// 3x3 stride-1 pad-1 patches of a few-channel channels_last image as GEMM rows: cols[p, (ky*3+kx)*Cin + ci] =
// x[b, h+ky-1, w+kx-1, ci] (zero outside the image and in the pad columns up to Kpad).  The stem convolution
// (stem_head.py:23-32, Cin = 3) then runs -- forward and weight gradient -- on the tcgen05 GEMM with K = 32: the image is
// 6 MB at batch 1024, the patches 67 MB, against 134 MB of output.  Thread = (pixel, 8 patch columns).
// ------------------------------------------------------------------------------------------------
template <typename T, int CIN>
__global__ void im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ cols, int H, int W, unsigned npix) {
  constexpr int KP = (9 * CIN + 7) / 8 * 8;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const unsigned w = p % (unsigned)W, t = p / (unsigned)W;
    const unsigned h = t % (unsigned)H;
    const T* centre = x + (size_t)p * CIN;
    float v[KP];
#pragma unroll
    for (int k = 9 * CIN; k < KP; ++k) v[k] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const bool in = (unsigned)((int)h + dy) < (unsigned)H && (unsigned)((int)w + dx) < (unsigned)W;
      const T* src = centre + (dy * W + dx) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) v[tap * CIN + ci] = in ? ld1(src + ci) : 0.f;
    }
    T* dst = cols + (size_t)p * KP;
#pragma unroll
    for (int g = 0; g < KP / 8; ++g) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[g * 8 + j];
      st8(dst + g * 8, o);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// The Downsample convolution (downsampling.py:41-47: 3x3, stride 2, pad 1, C -> 2C) as patches x tcgen05 GEMM.
// Patches of a channels_last image whose channel count is a multiple of 8: cols[(b, oy, ox), tap*Cin + c] =
// x[b, oy*s - 1 + ky, ox*s - 1 + kx, c].  Thread = (output pixel, 8 channels): the nine 16-byte loads are issued before the
// nine stores (144 bytes in flight per thread), Cin/8 neighbouring threads move one contiguous Cin-wide run on both sides.
// The input gradient is the reverse gather: dcols = dpre x W2 comes from the GEMM, and every input pixel sums the (at
// most four at stride 2, nine at stride 1) patch entries that were copies of it, in fp32 and in a fixed order.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void st_raw8(T* p, const Raw8<T>& r);
template <> __device__ __forceinline__ void st_raw8<float>(float* p, const Raw8<float>& r) {
  *reinterpret_cast<float4*>(p) = r.a;
  *reinterpret_cast<float4*>(p + 4) = r.b;
}
template <> __device__ __forceinline__ void st_raw8<bf16>(bf16* p, const Raw8<bf16>& r) {
  *reinterpret_cast<uint4*>(p) = r.u;
}

template <typename T, int STRIDE>
__global__ void __launch_bounds__(256) im2col3x3_vec_kernel(const T* __restrict__ x, T* __restrict__ cols, int H, int W,
                                                            int Ho, int Wo, int cg, unsigned nitems) {
  const int Cin = cg * 8;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
    const unsigned c = it % (unsigned)cg, p = it / (unsigned)cg;
    const unsigned ox = p % (unsigned)Wo, q = p / (unsigned)Wo;
    const unsigned oy = q % (unsigned)Ho, b = q / (unsigned)Ho;
    const int iy0 = (int)oy * STRIDE - 1, ix0 = (int)ox * STRIDE - 1;
    const T* img = x + (size_t)b * H * W * Cin + c * 8;
    Raw8<T> r[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = iy0 + tap / 3, ix = ix0 + tap % 3;
      raw8_zero(r[tap]);
      if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) ld_raw8(img + ((size_t)iy * W + ix) * Cin, r[tap]);
    }
    T* dst = cols + (size_t)p * 9 * Cin + c * 8;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) st_raw8(dst + tap * Cin, r[tap]);
  }
}

template <typename T, int STRIDE>
__global__ void __launch_bounds__(256) col2im3x3_vec_kernel(const T* __restrict__ dcols, T* __restrict__ dx, int H, int W,
                                                            int Ho, int Wo, int cg, unsigned nitems) {
  const int Cin = cg * 8;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
    const unsigned c = it % (unsigned)cg, p = it / (unsigned)cg;
    const unsigned ix = p % (unsigned)W, q = p / (unsigned)W;
    const unsigned iy = q % (unsigned)H, b = q / (unsigned)H;
    const T* base = dcols + (size_t)b * Ho * Wo * 9 * Cin + c * 8;
    Raw8<T> r[9];
    bool on[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      // the output position whose patch holds this pixel at (ky, kx): o*s - 1 + k = i
      const int ty = (int)iy + 1 - tap / 3, tx = (int)ix + 1 - tap % 3;
      const int oy = ty / STRIDE, oxx = tx / STRIDE;
      on[tap] = ty >= 0 && tx >= 0 && ty % STRIDE == 0 && tx % STRIDE == 0 && oy < Ho && oxx < Wo;
      if (on[tap]) ld_raw8(base + ((size_t)oy * Wo + oxx) * 9 * Cin + tap * Cin, r[tap]);
    }
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (on[tap]) {
        float v[8];
        cvt_raw8(r[tap], v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[i];
      }
    }
    st8(dx + (size_t)p * Cin + c * 8, acc);
  }
}

// stride 2, even H and W: thread = (2 x 2 input quad, 8 channels).  The quad (2j+py, 2i+px) collects exactly nine patch
// entries -- taps (1,1) (1,2) (2,1) (2,2) of output (j, i), (1,0) (2,0) of (j, i+1), (0,1) (0,2) of (j+1, i) and (0,0)
// of (j+1, i+1) -- so every thread has nine independent 16-byte loads in flight and no lane diverges on parity.
template <typename T>
__global__ void __launch_bounds__(256) col2im3x3_s2_quad_kernel(const T* __restrict__ dcols, T* __restrict__ dx, int Ho,
                                                                int Wo, int cg, unsigned nitems) {
  const int Cin = cg * 8;
  const size_t rowlen = (size_t)9 * Cin;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += gridDim.x * blockDim.x) {
    const unsigned c = it % (unsigned)cg, q = it / (unsigned)cg;  // q = (b*Ho + j)*Wo + i: the output pixel (j, i)
    const unsigned i = q % (unsigned)Wo, bj = q / (unsigned)Wo;
    const unsigned j = bj % (unsigned)Ho;
    const bool right = i + 1 < (unsigned)Wo, down = j + 1 < (unsigned)Ho;
    const T* r00 = dcols + (size_t)q * rowlen + c * 8;
    const T* r01 = r00 + rowlen;
    const T* r10 = r00 + (size_t)Wo * rowlen;
    const T* r11 = r10 + rowlen;
    Raw8<T> t4, t5, t7, t8, t3, t6, t1, t2, t0;
    ld_raw8(r00 + 4 * Cin, t4); ld_raw8(r00 + 5 * Cin, t5); ld_raw8(r00 + 7 * Cin, t7); ld_raw8(r00 + 8 * Cin, t8);
    raw8_zero(t3); raw8_zero(t6); raw8_zero(t1); raw8_zero(t2); raw8_zero(t0);
    if (right) { ld_raw8(r01 + 3 * Cin, t3); ld_raw8(r01 + 6 * Cin, t6); }
    if (down) { ld_raw8(r10 + 1 * Cin, t1); ld_raw8(r10 + 2 * Cin, t2); }
    if (right && down) ld_raw8(r11, t0);
    float a[8], b[8], o[8];
    // (2j, 2i)
    T* d00 = dx + (((size_t)bj * 2) * (2 * Wo) + 2 * i) * Cin + c * 8;
    cvt_raw8(t4, o);
    st8(d00, o);
    // (2j, 2i+1)
    cvt_raw8(t5, a); cvt_raw8(t3, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = a[k] + b[k];
    st8(d00 + Cin, o);
    // (2j+1, 2i)
    T* d10 = d00 + (size_t)2 * Wo * Cin;
    cvt_raw8(t7, a); cvt_raw8(t1, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = a[k] + b[k];
    st8(d10, o);
    // (2j+1, 2i+1)
    cvt_raw8(t8, a); cvt_raw8(t6, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = a[k] + b[k];
    cvt_raw8(t2, a); cvt_raw8(t0, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (o[k] + a[k]) + b[k];
    st8(d10 + Cin, o);
  }
}

}  // namespace

extern "C" int ogv_im2col3x3(const void* x, void* cols, int B, int H, int W, int Cin, int Kpad, int dtype, void* stream) {
  OGV_REQUIRE(x && cols && B > 0 && H > 0 && W > 0, "im2col3x3: bad args");
  OGV_REQUIRE(Cin >= 1 && Cin <= 7, "im2col3x3: Cin=%d (1..7: the patch row must fit one 64-wide K tile)", Cin);
  OGV_REQUIRE(Kpad == (9 * Cin + 7) / 8 * 8, "im2col3x3: Kpad=%d must be 9*Cin=%d rounded up to 8", Kpad, 9 * Cin);
  const long long npix = (long long)B * H * W;
  OGV_REQUIRE(npix < 0x7fffffffLL, "im2col3x3: too many pixels");
  long long blocks = (npix + 255) / 256;
  const long long cap = (long long)ogv_num_sms() * 16;
  if (blocks > cap) blocks = cap;
#define OGV_IM2COL_CASE(N)                                                                                         \
  case N:                                                                                                          \
    im2col3x3_kernel<T, N><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(x),      \
                                                                               reinterpret_cast<T*>(cols), H, W,   \
                                                                               (unsigned)npix);                    \
    break;
  OGV_DISPATCH_DTYPE(dtype, T, {
    switch (Cin) {
      OGV_IM2COL_CASE(1) OGV_IM2COL_CASE(2) OGV_IM2COL_CASE(3) OGV_IM2COL_CASE(4)
      OGV_IM2COL_CASE(5) OGV_IM2COL_CASE(6) OGV_IM2COL_CASE(7)
    }
    return ogv_check_launch("im2col3x3");
  });
#undef OGV_IM2COL_CASE
}

extern "C" int ogv_im2col3x3_vec(const void* x, void* cols, int B, int H, int W, int Cin, int stride, int dtype,
                                 void* stream) {
  OGV_REQUIRE(x && cols && B > 0 && H > 0 && W > 0, "im2col3x3_vec: bad args");
  OGV_REQUIRE(Cin > 0 && Cin % 8 == 0, "im2col3x3_vec: Cin=%d must be a multiple of 8", Cin);
  OGV_REQUIRE(stride == 1 || stride == 2, "im2col3x3_vec: stride=%d (1 or 2)", stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long nitems = (long long)B * Ho * Wo * (Cin / 8);
  OGV_REQUIRE(nitems < 0x7fffffffLL, "im2col3x3_vec: too many items");
  long long blocks = (nitems + 255) / 256;
  const long long cap = (long long)ogv_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  OGV_DISPATCH_DTYPE(dtype, T, {
    if (stride == 1)
      im2col3x3_vec_kernel<T, 1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(x), reinterpret_cast<T*>(cols), H, W, Ho, Wo, Cin / 8, (unsigned)nitems);
    else
      im2col3x3_vec_kernel<T, 2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(x), reinterpret_cast<T*>(cols), H, W, Ho, Wo, Cin / 8, (unsigned)nitems);
    return ogv_check_launch("im2col3x3_vec");
  });
}

extern "C" int ogv_col2im3x3_vec(const void* dcols, void* dx, int B, int H, int W, int Cin, int stride, int dtype,
                                 void* stream) {
  OGV_REQUIRE(dcols && dx && B > 0 && H > 0 && W > 0, "col2im3x3_vec: bad args");
  OGV_REQUIRE(Cin > 0 && Cin % 8 == 0, "col2im3x3_vec: Cin=%d must be a multiple of 8", Cin);
  OGV_REQUIRE(stride == 1 || stride == 2, "col2im3x3_vec: stride=%d (1 or 2)", stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long nitems = (long long)B * H * W * (Cin / 8);
  OGV_REQUIRE(nitems < 0x7fffffffLL, "col2im3x3_vec: too many items");
  long long blocks = (nitems + 255) / 256;
  const long long cap = (long long)ogv_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (stride == 2 && H % 2 == 0 && W % 2 == 0) {
    const long long nq = nitems / 4;
    long long qb = (nq + 255) / 256;
    if (qb > cap) qb = cap;
    OGV_DISPATCH_DTYPE(dtype, T, {
      col2im3x3_s2_quad_kernel<T><<<(unsigned)qb, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dcols), reinterpret_cast<T*>(dx), Ho, Wo, Cin / 8, (unsigned)nq);
      return ogv_check_launch("col2im3x3_vec");
    });
  }
  OGV_DISPATCH_DTYPE(dtype, T, {
    if (stride == 1)
      col2im3x3_vec_kernel<T, 1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dcols), reinterpret_cast<T*>(dx), H, W, Ho, Wo, Cin / 8, (unsigned)nitems);
    else
      col2im3x3_vec_kernel<T, 2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
          reinterpret_cast<const T*>(dcols), reinterpret_cast<T*>(dx), H, W, Ho, Wo, Cin / 8, (unsigned)nitems);
    return ogv_check_launch("col2im3x3_vec");
  });
}

extern "C" int ogv_nchw_to_nhwc(const void* src, void* dst, int B, int C, int HW, int dtype, void* stream) {
  OGV_REQUIRE(src && dst && B >= 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad args");
  OGV_DISPATCH_DTYPE(dtype, T, return transpose_batched<T>(src, dst, B, C, HW, (cudaStream_t)stream));
}
extern "C" int ogv_nhwc_to_nchw(const void* src, void* dst, int B, int C, int HW, int dtype, void* stream) {
  OGV_REQUIRE(src && dst && B >= 0 && C > 0 && HW > 0, "nhwc_to_nchw: bad args");
  OGV_DISPATCH_DTYPE(dtype, T, return transpose_batched<T>(src, dst, B, HW, C, (cudaStream_t)stream));
}

extern "C" int ogv_cast_transpose(const float* src, void* dst, long long ld_dst, void* dst_t, long long ld_dst_t,
                                  int rows, int cols, int dtype, void* stream) {
  OGV_REQUIRE(src && rows > 0 && cols > 0 && (dst || dst_t), "cast_transpose: bad args");
  dim3 grid(ogv_ceil_div(cols, 32), ogv_ceil_div(rows, 32));
  OGV_DISPATCH_DTYPE(dtype, T, {
    cast_transpose_kernel<T><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        src, reinterpret_cast<T*>(dst), ld_dst, reinterpret_cast<T*>(dst_t), ld_dst_t, rows, cols);
    return ogv_check_launch("cast_transpose");
  });
}

extern "C" int ogv_cast_batch(const ogv_cast_item* items, int n_items, int total_tiles, void* stream) {
  if (n_items == 0 || total_tiles == 0) return OGV_OK;
  OGV_REQUIRE(items && n_items > 0 && total_tiles > 0, "cast_batch: bad args");
  const int cap = ogv_num_sms() * 8;
  cast_batch_kernel<<<total_tiles < cap ? total_tiles : cap, dim3(32, 8), 0, (cudaStream_t)stream>>>(items, n_items,
                                                                                                  total_tiles);
  return ogv_check_launch("cast_batch");
}

extern "C" int ogv_split3(const float* src, long long ld_src, void* dst, long long ld_dst, long long plane_stride,
                          long long rows, int cols, int pattern, void* stream) {
  if (rows == 0 || cols == 0) return OGV_OK;
  OGV_REQUIRE(src && dst && (pattern == 0 || pattern == 1), "split3: bad args");
  OGV_REQUIRE(cols % 8 == 0 && ld_src % 4 == 0 && ld_dst % 8 == 0 && plane_stride % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "split3: rows must be 16-byte aligned and cols a multiple of 8");
  const long long n = rows * (cols / 8);
  OGV_REQUIRE(n < 0xffffffffLL, "split3: tensor too large");
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)ogv_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  split3_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, reinterpret_cast<bf16*>(dst), ld_dst,
                                                              plane_stride, rows, cols, pattern);
  return ogv_check_launch("split3");
}

extern "C" int ogv_rowscale(const void* x, const float* scale, void* y, long long rows, int cols,
                            int rows_per_scale, int dtype, void* stream) {
  OGV_REQUIRE(x && y && scale && cols % 8 == 0 && rows_per_scale > 0, "rowscale: bad args (cols %% 8 == 0)");
  long long nvec = rows * (cols / 8);
  if (nvec == 0) return OGV_OK;
  int grid = (int)((nvec + 255) / 256 < ogv_num_sms() * 16 ? (nvec + 255) / 256 : ogv_num_sms() * 16);
  OGV_DISPATCH_DTYPE(dtype, T, {
    rowscale_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(x), scale,
                                                               reinterpret_cast<T*>(y), nvec, cols / 8,
                                                               rows_per_scale);
    return ogv_check_launch("rowscale");
  });
}

extern "C" int ogv_rowscale_colsum(const void* x, const float* scale, void* y, float* out, long long rows, int cols,
                                   int rows_per_scale, int dtype, void* stream) {
  if (rows == 0 || cols == 0) return OGV_OK;
  OGV_REQUIRE(x && scale && out && cols % 8 == 0 && rows_per_scale > 0, "rowscale_colsum: bad args (cols %% 8 == 0)");  // y may be null: reduction only
  OGV_REQUIRE(rows < 0x7fffffffLL, "rowscale_colsum: too many rows");
  OGV_DISPATCH_DTYPE(dtype, T, {
    ColReduceCfg cfg;
    if (!colreduce_config(rows, cols / 8, &cfg, rowscale_colsum_kernel<T>)) { ogv_set_error("rowscale_colsum: cols=%d too wide", cols); return OGV_ERR_UNSUPPORTED; }
    rowscale_colsum_kernel<T><<<cfg.grid, cfg.block, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const T*>(x), scale, reinterpret_cast<T*>(y), out, rows, cols / 8, (unsigned)rows_per_scale);
    return ogv_check_launch("rowscale_colsum");
  });
}

extern "C" int ogv_mul_dact(const void* a, const void* pre, void* out, long long n, int act, int dtype,
                            void* stream) {
  if (n == 0) return OGV_OK;
  OGV_REQUIRE(a && pre && out, "mul_dact: null");
  int grid = (int)((n + 255) / 256 < ogv_num_sms() * 16 ? (n + 255) / 256 : ogv_num_sms() * 16);
  OGV_DISPATCH_DTYPE(dtype, T, {
    mul_dact_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(a),
                                                               reinterpret_cast<const T*>(pre),
                                                               reinterpret_cast<T*>(out), n, act);
    return ogv_check_launch("mul_dact");
  });
}

extern "C" int ogv_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream) {
  if (n == 0) return OGV_OK;
  OGV_REQUIRE(a && b && y, "add: null");
  int grid = (int)((n + 255) / 256 < ogv_num_sms() * 16 ? (n + 255) / 256 : ogv_num_sms() * 16);
  OGV_DISPATCH_DTYPE(dtype, T, {
    add_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(a),
                                                          reinterpret_cast<const T*>(b), reinterpret_cast<T*>(y), n);
    return ogv_check_launch("add");
  });
}

extern "C" int ogv_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, float bias_c1, float bias_c2, float grad_scale,
                         void* stream) {
  if (n == 0) return OGV_OK;
  OGV_REQUIRE(p && g && m && v, "adamw: null");
  int grid = (int)((n + 255) / 256 < ogv_num_sms() * 16 ? (n + 255) / 256 : ogv_num_sms() * 16);
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bias_c1,
                                                       bias_c2, grad_scale);
  return ogv_check_launch("adamw");
}

extern "C" int ogv_colsum(const void* x, long long ld, float* out, long long M, int N, int dtype, void* stream) {
  if (M == 0 || N == 0) return OGV_OK;
  OGV_REQUIRE(x && out && N % 8 == 0 && ld % 8 == 0, "colsum: N and ld must be multiples of 8");
  OGV_DISPATCH_DTYPE(dtype, T, {
    ColReduceCfg cfg;
    if (!colreduce_config(M, N / 8, &cfg, colsum_kernel<T>)) { ogv_set_error("colsum: N=%d too wide", N); return OGV_ERR_UNSUPPORTED; }
    colsum_kernel<T><<<cfg.grid, cfg.block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(x), ld, out, M,
                                                                         N / 8);
    return ogv_check_launch("colsum");
  });
}

extern "C" int ogv_colstats(const void* x, long long ld, float* sum, float* sumsq, long long M, int N, int dtype,
                            void* stream) {
  if (M == 0 || N == 0) return OGV_OK;
  OGV_REQUIRE(x && sum && sumsq && N % 8 == 0 && ld % 8 == 0, "colstats: N and ld must be multiples of 8");
  OGV_DISPATCH_DTYPE(dtype, T, {
    ColReduceCfg cfg;
    if (!colreduce_config(M, N / 8, &cfg, colstats_kernel<T>)) { ogv_set_error("colstats: N=%d too wide", N); return OGV_ERR_UNSUPPORTED; }
    colstats_kernel<T><<<cfg.grid, cfg.block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(x), ld, sum,
                                                                           sumsq, M, N / 8);
    return ogv_check_launch("colstats");
  });
}
