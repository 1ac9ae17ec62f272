// fp32-accumulate SIMT GEMM with generic operand strides.  This is the exact-fp32 engine used for
// the fp32 parity configuration (BASELINE config 1) and for the tiny SE / classifier-sized
// products; the bf16 training path runs on the tcgen05 engine (ogv_gemm_tc.cu).
#include "ogv_gemm.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename TI, typename TO>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const TI* __restrict__ A, long long a_rs, long long a_cs,
                                                       const TI* __restrict__ B, long long b_rs, long long b_cs,
                                                       int K, int k_per_split, GemmEpi e) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const bool a_kmajor = (a_cs == 1), b_kmajor = (b_cs == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int el = tid + i * NT;
      int r, k;
      if (a_kmajor) { r = el / BK; k = el % BK; } else { r = el % BM; k = el / BM; }
      int gm = m0 + r, gk = k0 + k;
      float v = 0.f;
      if (gm < e.M && gk < k_end) v = ld1(A + (long long)gm * a_rs + (long long)gk * a_cs);
      As[k][r] = v;
      if (b_kmajor) { r = el / BK; k = el % BK; } else { r = el % BN; k = el / BN; }
      int gn = n0 + r;
      gk = k0 + k;
      v = 0.f;
      if (gn < e.N && gk < k_end) v = ld1(B + (long long)gn * b_rs + (long long)gk * b_cs);
      Bs[k][r] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= e.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= e.N) continue;
      float s = epi_scalar<TO>(e, m, n, acc[i][j]);
      cs[j] += s;
      cq[j] += s * s;
    }
  }
  if (e.col_sum) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < e.N) {
        atomicAdd(e.col_sum + n, cs[j]);
        if (e.col_sumsq) atomicAdd(e.col_sumsq + n, cq[j]);
      }
    }
  }
}

template <typename TI, typename TO>
int launch(const ogv_gemm_args& a, cudaStream_t stream) {
  GemmEpi e = make_epi(a);
  int split = a.split_k > 0 ? a.split_k : 1;
  int k_per_split = ogv_ceil_div(ogv_ceil_div(a.K, split), BK) * BK;
  split = ogv_ceil_div(a.K, k_per_split);
  dim3 grid(ogv_ceil_div(a.M, BM), ogv_ceil_div(a.N, BN), split);
  gemm_simt_kernel<TI, TO><<<grid, NT, 0, stream>>>(reinterpret_cast<const TI*>(a.A), a.a_rs, a.a_cs,
                                                    reinterpret_cast<const TI*>(a.B), a.b_rs, a.b_cs, a.K,
                                                    k_per_split, e);
  return ogv_check_launch("gemm_simt");
}

}  // namespace

int ogv_gemm_simt(const ogv_gemm_args& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return OGV_OK;
  if (a.in_dtype == OGV_F32 && a.out_dtype == OGV_F32) return launch<float, float>(a, stream);
  if (a.in_dtype == OGV_BF16 && a.out_dtype == OGV_BF16) return launch<bf16, bf16>(a, stream);
  if (a.in_dtype == OGV_BF16 && a.out_dtype == OGV_F32) return launch<bf16, float>(a, stream);
  if (a.in_dtype == OGV_F32 && a.out_dtype == OGV_BF16) return launch<float, bf16>(a, stream);
  ogv_set_error("gemm_simt: bad dtype codes %d/%d", a.in_dtype, a.out_dtype);
  return OGV_ERR_ARG;
}
