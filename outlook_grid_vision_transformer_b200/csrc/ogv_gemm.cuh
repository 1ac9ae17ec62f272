// Shared epilogue for the pointwise GEMM kernels (SIMT fp32 and tcgen05 bf16).
#pragma once
#include "ogv_common.cuh"
#include "../../include/ogv.h"

struct GemmEpi {
  void* D;
  long long ldd;
  const float* bias;
  void* pre_out;
  long long ld_pre;
  int act;
  const void* dact_src;
  long long ld_dact;
  int dact;
  const float* row_scale;
  int rows_per_scale;
  const void* residual;
  long long ld_res;
  int accumulate;
  float* col_sum;
  float* col_sumsq;
  int pre_out_grad;
  int M, N;
};

static inline GemmEpi make_epi(const ogv_gemm_args& a) {
  GemmEpi e;
  e.D = a.D; e.ldd = a.ldd; e.bias = a.bias; e.pre_out = a.pre_out; e.ld_pre = a.ld_pre; e.act = a.act;
  e.dact_src = a.dact_src; e.ld_dact = a.ld_dact; e.dact = a.dact; e.row_scale = a.row_scale;
  e.rows_per_scale = a.rows_per_scale > 0 ? a.rows_per_scale : 1; e.residual = a.residual; e.ld_res = a.ld_res;
  e.accumulate = a.accumulate; e.col_sum = a.col_sum; e.col_sumsq = a.col_sumsq; e.pre_out_grad = a.pre_out_grad; e.M = a.M; e.N = a.N;
  return e;
}

// Scalar epilogue for one element; returns the value that was stored (after rounding to TO).
template <typename TO>
__device__ __forceinline__ float epi_scalar(const GemmEpi& e, int m, int n, float v) {
  if (e.bias) v += e.bias[n];
  if (e.pre_out)
    st1(reinterpret_cast<TO*>(e.pre_out) + (long long)m * e.ld_pre + n, e.pre_out_grad ? act_grad(e.act, v) : v);
  v = act_apply(e.act, v);
  if (e.dact_src) v *= act_grad(e.dact, ld1(reinterpret_cast<const TO*>(e.dact_src) + (long long)m * e.ld_dact + n));
  if (e.row_scale) v *= e.row_scale[m / e.rows_per_scale];
  if (e.residual) v += ld1(reinterpret_cast<const TO*>(e.residual) + (long long)m * e.ld_res + n);
  if (e.accumulate) {
    atomicAdd(reinterpret_cast<float*>(e.D) + (long long)m * e.ldd + n, v);
    return v;
  }
  TO* d = reinterpret_cast<TO*>(e.D) + (long long)m * e.ldd + n;
  st1(d, v);
  return round_to<TO>(v);
}

// One operand of a tcgen05 GEMM as the never-materialised 3x3 / pad 1 patch matrix of a channels_last bf16 image
// (ogv_gemm_tc.cu: ConvView).  which = 1: the A operand of a forward product (args.A = x, a_rs = 9*Cin, a_cs = 1,
// M = B*Ho*Wo, K = 9*Cin); which = 2: the B operand of a weight-gradient product (args.B = x, b_rs = 1, b_cs = 9*Cin,
// N = 9*Cin, K = B*Ho*Wo).
struct ogv_conv_view {
  int which;
  const void* x;
  int B, H, W, Cin, stride;
};
bool ogv_conv_view_supported(const ogv_conv_view& c);
int ogv_gemm_tc_conv(const ogv_gemm_args& a, const ogv_conv_view& conv, cudaStream_t stream);
int ogv_gemm_simt(const ogv_gemm_args& a, cudaStream_t stream);
int ogv_gemm_tc(const ogv_gemm_args& a, cudaStream_t stream);
bool ogv_gemm_tc_supported(const ogv_gemm_args& a, const char** why);
