// Common device/host helpers for the OutGridBlock sm_100a kernels.
// Activations are stored "rows x channels" (NHWC flattened: row m = (b*H + h)*W + w),
// either fp32 or bf16; all arithmetic is fp32.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

#define OGV_OK 0
#define OGV_ERR_ARG (-1)
#define OGV_ERR_CUDA (-2)
#define OGV_ERR_UNSUPPORTED (-3)

#define OGV_F32 0
#define OGV_BF16 1

#define OGV_ACT_NONE 0
#define OGV_ACT_GELU 1
#define OGV_ACT_SILU 2
#define OGV_ACT_SIGMOID 3
#define OGV_ACT_RELU 4
#define OGV_ACT_MUL 5

void ogv_set_error(const char* fmt, ...);
int ogv_check_launch(const char* what);
int ogv_num_sms();

#define OGV_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ogv_set_error(__VA_ARGS__);              \
      return OGV_ERR_ARG;                      \
    }                                          \
  } while (0)

// Dispatch on the activation dtype code.
#define OGV_DISPATCH_DTYPE(dtype, T, ...)                       \
  do {                                                          \
    if ((dtype) == OGV_F32) {                                   \
      typedef float T;                                          \
      __VA_ARGS__;                                              \
    } else if ((dtype) == OGV_BF16) {                           \
      typedef bf16 T;                                           \
      __VA_ARGS__;                                              \
    } else {                                                    \
      ogv_set_error("unsupported dtype code %d", (int)(dtype)); \
      return OGV_ERR_ARG;                                       \
    }                                                           \
  } while (0)

// ---------------------------------------------------------------------------------------------
// scalar / vector load-store in fp32 registers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// value as it will read back after being stored as T
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// 8 consecutive elements (16-byte aligned for bf16, 32-byte for fp32).
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// 8 elements kept in their storage format (4 registers for bf16): streaming kernels issue the loads of several
// rows up front and convert a row only when they get to it, so bytes in flight cost half the registers.
template <typename T> struct Raw8;
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<bf16> { uint4 u; };
__device__ __forceinline__ void raw8_zero(Raw8<float>& r) { r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; }
__device__ __forceinline__ void raw8_zero(Raw8<bf16>& r) { r.u = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void ld_raw8(const float* p, Raw8<float>& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void ld_raw8(const bf16* p, Raw8<bf16>& r) { r.u = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void cvt_raw8(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
  v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void cvt_raw8(const Raw8<bf16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// 4 elements in storage format (8 bytes of bf16 / 16 bytes of fp32)
template <typename T> struct Raw4;
template <> struct Raw4<float> { float4 a; };
template <> struct Raw4<bf16> { uint2 u; };
__device__ __forceinline__ void raw4_zero(Raw4<float>& r) { r.a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void raw4_zero(Raw4<bf16>& r) { r.u = make_uint2(0u, 0u); }
__device__ __forceinline__ void ld_raw4(const float* p, Raw4<float>& r) { r.a = *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void ld_raw4(const bf16* p, Raw4<bf16>& r) { r.u = *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void cvt_raw4(const Raw4<float>& r, float (&v)[4]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
}
__device__ __forceinline__ void cvt_raw4(const Raw4<bf16>& r, float (&v)[4]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
  const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
  v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}

// VEC-wide generic versions (VEC in {1,2,4,8}); alignment = VEC*sizeof(T).
template <int VEC, typename T>
__device__ __forceinline__ void ldv(const T* p, float (&v)[VEC]) {
  if constexpr (VEC == 8) {
    ld8(p, v);
  } else if constexpr (VEC == 4 && sizeof(T) == 4) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else if constexpr (VEC == 4 && sizeof(T) == 2) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
  } else if constexpr (VEC == 2 && sizeof(T) == 2) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    v[0] = f.x; v[1] = f.y;
  } else if constexpr (VEC == 2 && sizeof(T) == 4) {
    const float2 f = *reinterpret_cast<const float2*>(p);
    v[0] = f.x; v[1] = f.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = ld1(p + i);
  }
}
template <int VEC, typename T>
__device__ __forceinline__ void stv(T* p, const float (&v)[VEC]) {
  if constexpr (VEC == 8) {
    st8(p, v);
  } else if constexpr (VEC == 4 && sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 4 && sizeof(T) == 2) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  } else if constexpr (VEC == 2 && sizeof(T) == 2) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
  } else if constexpr (VEC == 2 && sizeof(T) == 4) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) st1(p + i, v[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 -- two IEEE round-to-nearest fp32 operations per
// issue slot).  Results are bit-identical to the scalar forms; the gain is issue bandwidth in the
// instruction-bound stencil and epilogue loops.  A pair lives in one 64-bit register (element 0 in the
// low half = the lower address of a float2), so 8- and 16-byte loads deliver pairs without any packing.
// -DOGV_PACKED_F32=0 builds the scalar forms (A/B measurements).
// ---------------------------------------------------------------------------------------------
#ifndef OGV_PACKED_F32
#define OGV_PACKED_F32 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
#if OGV_PACKED_F32
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
#else
  float a0, a1, b0, b1, c0, c1;
  unpk2(a, a0, a1); unpk2(b, b0, b1); unpk2(c, c0, c1);
  return pk2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
#endif
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
#if OGV_PACKED_F32
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
#else
  float a0, a1, b0, b1;
  unpk2(a, a0, a1); unpk2(b, b0, b1);
  return pk2(a0 * b0, a1 * b1);
#endif
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
#if OGV_PACKED_F32
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
#else
  float a0, a1, b0, b1;
  unpk2(a, a0, a1); unpk2(b, b0, b1);
  return pk2(a0 + b0, a1 + b1);
#endif
}

// ---------------------------------------------------------------------------------------------
// activations (exact erf GELU, SiLU, sigmoid) and their derivatives w.r.t. the pre-activation
// ---------------------------------------------------------------------------------------------
// Transcendentals are MUFU-throughput bound on the wide (4C) tensors, so the fast hardware paths are
// used: ex2.approx / rcp.approx (rel. error ~1e-6, far below both parity tolerances) and the
// Abramowitz-Stegun 7.1.26 rational erf (abs. error 1.5e-7) sharing ONE exponential between
// erf(x/sqrt2) and the Gaussian pdf of GELU'.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float dsilu_f(float x) {
  float s = sigmoid_f(x);
  return s * (1.f + x * (1.f - s));
}
// returns erf(x / sqrt(2)); *gauss = exp(-x^2 / 2)
__device__ __forceinline__ float erf_gauss(float x, float* gauss) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float e = __expf(-z * z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.f - p * t * e;
  *gauss = e;
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_f(float x) {
  float e;
  return 0.5f * x * (1.f + erf_gauss(x, &e));
}
__device__ __forceinline__ float dgelu_f(float x) {
  float e;
  const float cdf = 0.5f * (1.f + erf_gauss(x, &e));
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float act_apply(int act, float x) {
  switch (act) {
    case OGV_ACT_GELU: return gelu_f(x);
    case OGV_ACT_SILU: return silu_f(x);
    case OGV_ACT_SIGMOID: return sigmoid_f(x);
    case OGV_ACT_RELU: return x > 0.f ? x : 0.f;
    default: return x;
  }
}
__device__ __forceinline__ float act_grad(int act, float x) {
  switch (act) {
    case OGV_ACT_GELU: return dgelu_f(x);
    case OGV_ACT_SILU: return dsilu_f(x);
    case OGV_ACT_SIGMOID: { float s = sigmoid_f(x); return s * (1.f - s); }
    case OGV_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case OGV_ACT_MUL: return x;
    default: return 1.f;
  }
}

// ---------------------------------------------------------------------------------------------
// bf16-mode activations: ONE MUFU op (tanh.approx.f32, rel. error 2^-11) per element instead of two
// (ex2 + rcp) and half the ALU work.  Their absolute error (<= 5e-4 on O(1) values) sits 4-8x below
// the bf16 quantisation of the stored result, so they are used only when the tensor dtype is bf16;
// fp32 tensors keep the accurate forms above.
//   sigmoid(x) = 0.5 + 0.5 tanh(x/2)
//   erf(z)    ~= tanh(z (c1 + c3 z^2 + c5 z^4))   (least-squares fit on GELU, max |gelu err| 3e-5)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
// Derivatives feed cancellation-heavy reductions (BatchNorm gamma / beta gradients), where the x-correlated
// 2^-11 error of tanh.approx shows up as a bias: the backward kernels keep the 2-MUFU sigmoid (ex2 + rcp,
// rel. error ~1e-7) -- same ALU count, one more MUFU op.
__device__ __forceinline__ void silu_both_fast(float x, float* a, float* da) {
  const float s = sigmoid_f(x);
  const float v = x * s;
  *a = v;
  *da = fmaf(v, 1.f - s, s);
}
constexpr float kGeluC1 = 1.12777659f, kGeluC3 = 0.1047942f, kGeluC5 = -0.0020293f;
// The fit holds for |x| <= 7; beyond, its negative z^4 coefficient would turn the argument around (|x| > 10.5 gave
// gelu(x) = x/2 or gelu(-x) = -x).  z^2 is clamped at its |x| = 7 value, where tanh is saturated (|arg| > 12) for good.
constexpr float kGeluZ2Max = 24.5f;
__device__ __forceinline__ float gelu_fast(float x) {
  const float z2 = fminf(0.5f * x * x, kGeluZ2Max);
  const float p = x * 0.70710678118654752440f * fmaf(z2, fmaf(z2, kGeluC5, kGeluC3), kGeluC1);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(p), h);
}
__device__ __forceinline__ void gelu_both_fast(float x, float* a, float* da) {
  const float z2 = fminf(0.5f * x * x, kGeluZ2Max);
  const float p = x * 0.70710678118654752440f * fmaf(z2, fmaf(z2, kGeluC5, kGeluC3), kGeluC1);
  const float dp = 0.70710678118654752440f * fmaf(z2, fmaf(z2, 5.f * kGeluC5, 3.f * kGeluC3), kGeluC1);
  const float t = tanh_approx(p);
  const float cdf = fmaf(0.5f, t, 0.5f);
  *a = x * cdf;
  *da = fmaf(0.5f * x * dp, fmaf(-t, t, 1.f), cdf);
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float a, d;
  gelu_both_fast(x, &a, &d);
  return d;
}
__device__ __forceinline__ float dsilu_fast(float x) {
  float a, d;
  silu_both_fast(x, &a, &d);
  return d;
}

// activation and its derivative from ONE transcendental evaluation
__device__ __forceinline__ void act_both(int act, float x, float* a, float* da) {
  switch (act) {
    case OGV_ACT_GELU: {
      float e;
      const float cdf = 0.5f * (1.f + erf_gauss(x, &e));
      *a = x * cdf;
      *da = fmaf(x * 0.39894228040143267794f, e, cdf);
      return;
    }
    case OGV_ACT_SILU: {
      const float s = sigmoid_f(x);
      *a = x * s;
      *da = s * (1.f + x * (1.f - s));
      return;
    }
    case OGV_ACT_SIGMOID: {
      const float s = sigmoid_f(x);
      *a = s;
      *da = s * (1.f - s);
      return;
    }
    case OGV_ACT_RELU: *a = x > 0.f ? x : 0.f; *da = x > 0.f ? 1.f : 0.f; return;
    default: *a = x; *da = 1.f; return;
  }
}

// Compile-time activation selection: a run-time `switch` inside a per-element loop puts every
// element in its own basic block (no scheduling across elements, MUFU latency fully exposed), so
// the hot kernels are instantiated per activation and dispatched on the host.
template <int ACT, bool FAST = false> __device__ __forceinline__ float act_apply_t(float x) {
  if constexpr (ACT == OGV_ACT_GELU) return FAST ? gelu_fast(x) : gelu_f(x);
  else if constexpr (ACT == OGV_ACT_SILU) return FAST ? silu_fast(x) : silu_f(x);
  else if constexpr (ACT == OGV_ACT_SIGMOID) return FAST ? sigmoid_fast(x) : sigmoid_f(x);
  else if constexpr (ACT == OGV_ACT_RELU) return x > 0.f ? x : 0.f;
  else return x;
}
template <int ACT, bool FAST = false> __device__ __forceinline__ float act_grad_t(float x) {
  if constexpr (ACT == OGV_ACT_GELU) return FAST ? dgelu_fast(x) : dgelu_f(x);
  else if constexpr (ACT == OGV_ACT_SILU) return FAST ? dsilu_fast(x) : dsilu_f(x);
  else if constexpr (ACT == OGV_ACT_SIGMOID) { const float s = FAST ? sigmoid_fast(x) : sigmoid_f(x); return s * (1.f - s); }
  else if constexpr (ACT == OGV_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  else return 1.f;
}
template <int ACT, bool FAST = false> __device__ __forceinline__ void act_both_t(float x, float* a, float* da) {
  if constexpr (FAST && ACT == OGV_ACT_GELU) gelu_both_fast(x, a, da);
  else if constexpr (FAST && ACT == OGV_ACT_SILU) silu_both_fast(x, a, da);
  else act_both(ACT, x, a, da);
}
// FAST is selected by the tensor dtype
template <typename T> struct FastAct { static constexpr bool value = sizeof(T) == 2; };

// in-place activation / multiply-by-derivative over a small register array with the activation
// `switch` hoisted OUT of the element loop (one branch per N elements, straight-line code inside)
template <int N, bool FAST = false> __device__ __forceinline__ void act_apply_n(int act, float (&v)[N]) {
  switch (act) {
    case OGV_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = FAST ? gelu_fast(v[i]) : gelu_f(v[i]);
      break;
    case OGV_ACT_SILU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = FAST ? silu_fast(v[i]) : silu_f(v[i]);
      break;
    case OGV_ACT_SIGMOID:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = FAST ? sigmoid_fast(v[i]) : sigmoid_f(v[i]);
      break;
    case OGV_ACT_RELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
      break;
    default: break;
  }
}
template <int N, bool FAST = false> __device__ __forceinline__ void act_grad_mul_n(int act, float (&v)[N], const float (&src)[N]) {
  switch (act) {
    case OGV_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] *= FAST ? dgelu_fast(src[i]) : dgelu_f(src[i]);
      break;
    case OGV_ACT_SILU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] *= FAST ? dsilu_fast(src[i]) : dsilu_f(src[i]);
      break;
    case OGV_ACT_SIGMOID:
#pragma unroll
      for (int i = 0; i < N; ++i) { const float sg = FAST ? sigmoid_fast(src[i]) : sigmoid_f(src[i]); v[i] *= sg * (1.f - sg); }
      break;
    case OGV_ACT_RELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = src[i] > 0.f ? v[i] : 0.f;
      break;
    case OGV_ACT_MUL:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] *= src[i];
      break;
    default: break;
  }
}
// v <- act(v), d <- act'(v) with the switch hoisted out of the element loop
template <int N, bool FAST = false> __device__ __forceinline__ void act_both_n(int act, float (&v)[N], float (&d)[N]) {
  switch (act) {
    case OGV_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; ++i) { if (FAST) gelu_both_fast(v[i], &v[i], &d[i]); else act_both(OGV_ACT_GELU, v[i], &v[i], &d[i]); }
      break;
    case OGV_ACT_SILU:
#pragma unroll
      for (int i = 0; i < N; ++i) { if (FAST) silu_both_fast(v[i], &v[i], &d[i]); else act_both(OGV_ACT_SILU, v[i], &v[i], &d[i]); }
      break;
    default:
#pragma unroll
      for (int i = 0; i < N; ++i) act_both(act, v[i], &v[i], &d[i]);
      break;
  }
}

// ---- the same three helpers over packed pairs (v[i] = elements 2i, 2i+1): the bf16-mode GELU forms and the
// plain multiply run as FFMA2 / FMUL2 (half the issue slots, bit-identical to the scalar forms); everything
// else unpacks and goes through the scalar helpers. ----
__device__ __forceinline__ f32x2 splat2(float c) { return pk2(c, c); }
__device__ __forceinline__ f32x2 tanh_approx2(f32x2 p) {
  float a, b;
  unpk2(p, a, b);
  return pk2(tanh_approx(a), tanh_approx(b));
}
__device__ __forceinline__ f32x2 min2(f32x2 a, float c) {
  float lo, hi;
  unpk2(a, lo, hi);
  return pk2(fminf(lo, c), fminf(hi, c));
}
__device__ __forceinline__ f32x2 gelu_fast2(f32x2 x) {
  const f32x2 half = splat2(0.5f);
  const f32x2 z2 = min2(mul2(mul2(half, x), x), kGeluZ2Max);
  const f32x2 q = fma2(z2, fma2(z2, splat2(kGeluC5), splat2(kGeluC3)), splat2(kGeluC1));
  const f32x2 t = tanh_approx2(mul2(mul2(x, splat2(0.70710678118654752440f)), q));
  const f32x2 h = mul2(half, x);
  return fma2(h, t, h);
}
__device__ __forceinline__ void gelu_both_fast2(f32x2 x, f32x2* a, f32x2* da) {
  const f32x2 half = splat2(0.5f), rs2 = splat2(0.70710678118654752440f);
  const f32x2 h = mul2(half, x);
  const f32x2 z2 = min2(mul2(h, x), kGeluZ2Max);
  const f32x2 q = fma2(z2, fma2(z2, splat2(kGeluC5), splat2(kGeluC3)), splat2(kGeluC1));
  const f32x2 dp = mul2(rs2, fma2(z2, fma2(z2, splat2(5.f * kGeluC5), splat2(3.f * kGeluC3)), splat2(kGeluC1)));
  const f32x2 t = tanh_approx2(mul2(mul2(x, rs2), q));
  const f32x2 cdf = fma2(half, t, half);
  *a = mul2(x, cdf);
  const f32x2 sech2 = fma2(mul2(t, splat2(-1.f)), t, splat2(1.f));
  *da = fma2(mul2(h, dp), sech2, cdf);
}
template <int N2> __device__ __forceinline__ void unpack_n(const f32x2 (&v)[N2], float (&f)[2 * N2]) {
#pragma unroll
  for (int i = 0; i < N2; ++i) unpk2(v[i], f[2 * i], f[2 * i + 1]);
}
template <int N2> __device__ __forceinline__ void pack_n(const float (&f)[2 * N2], f32x2 (&v)[N2]) {
#pragma unroll
  for (int i = 0; i < N2; ++i) v[i] = pk2(f[2 * i], f[2 * i + 1]);
}
template <int N2, bool FAST = false> __device__ __forceinline__ void act_apply_p(int act, f32x2 (&v)[N2]) {
  if (act == OGV_ACT_NONE) return;
  if (FAST && act == OGV_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < N2; ++i) v[i] = gelu_fast2(v[i]);
    return;
  }
  float f[2 * N2];
  unpack_n<N2>(v, f);
  act_apply_n<2 * N2, FAST>(act, f);
  pack_n<N2>(f, v);
}
template <int N2, bool FAST = false> __device__ __forceinline__ void act_grad_mul_p(int act, f32x2 (&v)[N2], const f32x2 (&src)[N2]) {
  if (act == OGV_ACT_MUL) {
#pragma unroll
    for (int i = 0; i < N2; ++i) v[i] = mul2(v[i], src[i]);
    return;
  }
  if (FAST && act == OGV_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < N2; ++i) {
      f32x2 a, d;
      gelu_both_fast2(src[i], &a, &d);
      v[i] = mul2(v[i], d);
    }
    return;
  }
  float f[2 * N2], g[2 * N2];
  unpack_n<N2>(v, f);
  unpack_n<N2>(src, g);
  act_grad_mul_n<2 * N2, FAST>(act, f, g);
  pack_n<N2>(f, v);
}
template <int N2, bool FAST = false> __device__ __forceinline__ void act_both_p(int act, f32x2 (&v)[N2], f32x2 (&d)[N2]) {
  if (FAST && act == OGV_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < N2; ++i) gelu_both_fast2(v[i], &v[i], &d[i]);
    return;
  }
  float f[2 * N2], g[2 * N2];
  unpack_n<N2>(v, f);
  act_both_n<2 * N2, FAST>(act, f, g);
  pack_n<N2>(f, v);
  pack_n<N2>(g, d);
}

#define OGV_ACT_CASE_(code, ACT, ...) \
  case code: {                        \
    constexpr int ACT = code;         \
    __VA_ARGS__;                      \
  } break;
#define OGV_DISPATCH_ACT(act, ACT, ...)                            \
  switch (act) {                                                   \
    OGV_ACT_CASE_(OGV_ACT_NONE, ACT, __VA_ARGS__)                  \
    OGV_ACT_CASE_(OGV_ACT_GELU, ACT, __VA_ARGS__)                  \
    OGV_ACT_CASE_(OGV_ACT_SILU, ACT, __VA_ARGS__)                  \
    OGV_ACT_CASE_(OGV_ACT_SIGMOID, ACT, __VA_ARGS__)               \
    OGV_ACT_CASE_(OGV_ACT_RELU, ACT, __VA_ARGS__)                  \
    default:                                                       \
      ogv_set_error("unknown activation code %d", (int)(act));     \
      return OGV_ERR_ARG;                                          \
  }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ogv_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
