// Sync-free tail of the train step over FLAT fp32 arenas (one_epoch_train.py:98-166, warmup.py:4-59,
// metrics.py:7-24): global gradient norm, clip + AdamW with every step-dependent scalar (learning rate,
// bias corrections, 1/world) read from DEVICE memory so a captured CUDA graph keeps following the
// host-side schedule, the reference's non-finite guard evaluated on the device, and the loss /
// top-k counters accumulated on the device.
#include "ogv_common.cuh"

namespace {

constexpr int OPT_THREADS = 256;

// out[0] += sum g[i]^2.  float4 streaming, one atomic per CTA.
__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const float* __restrict__ g, long long n4, long long n,
                                                             float* __restrict__ out) {
  float acc = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)OPT_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * OPT_THREADS) {
    const float4 v = __ldg(g4 + i);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0)
    for (long long i = 4 * n4 + threadIdx.x; i < n; i += OPT_THREADS) acc = fmaf(g[i], g[i], acc);
  acc = warp_sum(acc);
  __shared__ float part[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < OPT_THREADS / 32 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// hyper (device, fp32): [0] lr  [1] 1-beta1^t  [2] 1-beta2^t  [3] grad_scale (1/world)  [4] max_norm (<= 0: no clip)
// One thread per 8-element granule (every parameter starts on a granule boundary); bit i of decay_bits says
// whether granule i takes weight decay (the reference's two param groups, warmup.py:4-26).
__global__ void __launch_bounds__(OPT_THREADS)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  const unsigned* __restrict__ decay_bits, long long granules, const float* __restrict__ hyper,
                  const float* __restrict__ gnorm_sq, const float* __restrict__ loss, float b1, float b2, float eps,
                  float wd, float* __restrict__ skipped) {
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2];
  float gs = hyper[3];
  bool skip = false;
  if (loss) skip = !isfinite(*loss);  // one_epoch_train.py:99-109: a non-finite loss skips the update
  if (gnorm_sq) {
    const float nsq = *gnorm_sq;
    skip = skip || !isfinite(nsq);
    const float max_norm = hyper[4];
    if (max_norm > 0.f) {  // torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6))
      const float norm = sqrtf(nsq) * gs;
      gs *= fminf(1.f, max_norm / (norm + 1e-6f));
    }
  }
  if (skip) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && skipped) *skipped += 1.f;
    return;
  }
  const float step = lr / bc1, rbc2 = rsqrtf(bc2);
  for (long long gi = blockIdx.x * (long long)OPT_THREADS + threadIdx.x; gi < granules;
       gi += (long long)gridDim.x * OPT_THREADS) {
    const bool decay = (decay_bits[gi >> 5] >> (gi & 31)) & 1u;
    const float keep = decay ? 1.f - lr * wd : 1.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long i = gi * 2 + h;
      const float4 g4 = reinterpret_cast<const float4*>(g)[i];
      float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
      const float gg[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
      float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mm[j] = fmaf(b1, mm[j], (1.f - b1) * gg[j]);
        vv[j] = fmaf(b2, vv[j], (1.f - b2) * gg[j] * gg[j]);
        const float denom = fmaf(sqrtf(vv[j]), rbc2, eps);
        pp[j] = fmaf(-step, __fdividef(mm[j], denom), pp[j] * keep);
      }
      reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
      reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
  }
}

// acc (fp32 [5]) += [ loss * B, top-1 hits, top-3 hits, top-5 hits, B ]   (metrics.py:7-24, one_epoch_train.py:155-166)
// one warp per sample: rank of the label's logit = number of strictly larger logits (+ equal ones at a lower index,
// which is how torch.topk breaks ties).
__global__ void __launch_bounds__(128) metrics_kernel(const float* __restrict__ logits, long long ld,
                                                       const long long* __restrict__ labels, int B, int K,
                                                       const float* __restrict__ loss, float* __restrict__ acc) {
  const int warp = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* row = logits + (long long)warp * ld;
  const long long y = labels[warp];
  int rank = 0;
  if (y >= 0 && y < K) {
    const float ly = row[y];
    for (int k = lane; k < K; k += 32) {
      const float l = row[k];
      rank += (l > ly) || (l == ly && k < y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  } else {
    rank = K;
  }
  if (lane == 0) {
    if (rank < 1) atomicAdd(acc + 1, 1.f);
    if (rank < 3) atomicAdd(acc + 2, 1.f);
    if (rank < 5) atomicAdd(acc + 3, 1.f);
    if (warp == 0) {
      if (loss) atomicAdd(acc + 0, *loss * (float)B);
      atomicAdd(acc + 4, (float)B);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cross entropy with label smoothing, mean over the batch (the criterion of the reference's loop:
// nn.CrossEntropyLoss(label_smoothing=eps), train_full_model.py:52 / one_epoch_train.py:95-97):
//   loss_i = (1 - eps) * (lse_i - x_i[y_i]) + eps * (lse_i - mean_k x_i[k]),   loss = mean_i loss_i
//   dx_i[k] = g / B * (softmax_i[k] - (1 - eps) * [k == y_i] - eps / K)
// A warp per row (K classes strided over the lanes, three passes over a row that sits in L1), the batch mean as one
// atomic per row; backward re-reads the logits and the saved log-sum-exp.  Replaces ~25 element-wise / reduction
// launches of the composed loss inside the captured step.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void xent_fwd_kernel(const T* __restrict__ x, long long ld, const long long* __restrict__ labels, int B, int K,
                                float smoothing, float* __restrict__ lse_out, float* __restrict__ loss) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= B) return;
  const T* xr = x + (long long)row * ld;
  float mx = -INFINITY, sum = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float v = ld1(xr + k);
    mx = fmaxf(mx, v);
    sum += v;
  }
  mx = warp_max(mx);
  sum = warp_sum(sum);
  float se = 0.f;
  for (int k = lane; k < K; k += 32) se += __expf(ld1(xr + k) - mx);
  se = warp_sum(se);
  if (lane == 0) {
    const float lse = mx + __logf(se);
    const long long y = labels[row];
    const float xy = (y >= 0 && y < K) ? ld1(xr + y) : 0.f;
    const float li = (1.f - smoothing) * (lse - xy) + smoothing * (lse - sum / (float)K);
    lse_out[row] = lse;
    atomicAdd(loss, li / (float)B);
  }
}

template <typename T>
__global__ void xent_bwd_kernel(const T* __restrict__ x, long long ld, const long long* __restrict__ labels,
                                const float* __restrict__ lse, const float* __restrict__ gout, int B, int K,
                                float smoothing, T* __restrict__ dx, long long ldd) {
  const long long n = (long long)B * K;
  const float g = (gout ? *gout : 1.f) / (float)B;
  const float off = smoothing / (float)K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / K), k = (int)(i - (long long)row * K);
    const float p = __expf(ld1(x + (long long)row * ld + k) - lse[row]);
    const float t = (labels[row] == k) ? (1.f - smoothing) : 0.f;
    st1(dx + (long long)row * ldd + k, g * (p - t - off));
  }
}

}  // namespace

extern "C" int ogv_xent_fwd(const void* logits, long long ld, const long long* labels, int B, int K, float smoothing,
                            int dtype, float* lse, float* loss, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(logits && labels && lse && loss && K > 0 && ld >= K, "xent_fwd: bad arguments");
  OGV_REQUIRE(smoothing >= 0.f && smoothing <= 1.f, "xent_fwd: label smoothing %f outside [0, 1]", (double)smoothing);
  OGV_DISPATCH_DTYPE(dtype, T, {
    xent_fwd_kernel<T><<<ogv_ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(logits), ld, labels,
                                                                              B, K, smoothing, lse, loss);
    return ogv_check_launch("xent_fwd");
  });
}

extern "C" int ogv_xent_bwd(const void* logits, long long ld, const long long* labels, const float* lse,
                            const float* gout, int B, int K, float smoothing, int dtype, void* dlogits, long long ldd,
                            void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(logits && labels && lse && dlogits && K > 0 && ld >= K && ldd >= K, "xent_bwd: bad arguments");
  const long long n = (long long)B * K;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)ogv_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  OGV_DISPATCH_DTYPE(dtype, T, {
    xent_bwd_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(logits), ld, labels, lse,
                                                                            gout, B, K, smoothing,
                                                                            reinterpret_cast<T*>(dlogits), ldd);
    return ogv_check_launch("xent_bwd");
  });
}

extern "C" int ogv_sumsq(const float* g, long long n, float* out, void* stream) {
  if (n == 0) return OGV_OK;
  OGV_REQUIRE(g && out, "sumsq: null");
  OGV_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: arena must be 16-byte aligned");
  const long long n4 = n / 4;
  long long want = (n4 + OPT_THREADS * 8 - 1) / (OPT_THREADS * 8);
  const int cap = ogv_num_sms() * 8;
  const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  sumsq_kernel<<<grid, OPT_THREADS, 0, (cudaStream_t)stream>>>(g, n4, n, out);
  return ogv_check_launch("sumsq");
}

extern "C" int ogv_adamw_flat(float* p, const float* g, float* m, float* v, const unsigned* decay_bits, long long n,
                              const float* hyper, const float* gnorm_sq, const float* loss, float beta1, float beta2,
                              float eps, float weight_decay, float* skipped, void* stream) {
  if (n == 0) return OGV_OK;
  OGV_REQUIRE(p && g && m && v && decay_bits && hyper, "adamw_flat: null");
  OGV_REQUIRE(n % 8 == 0, "adamw_flat: the arena length must be a multiple of 8 floats (granules)");
  for (const void* q : {(const void*)p, (const void*)g, (const void*)m, (const void*)v})
    OGV_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "adamw_flat: arenas must be 16-byte aligned");
  const long long granules = n / 8;
  long long want = (granules + OPT_THREADS - 1) / OPT_THREADS;
  const int cap = ogv_num_sms() * 8;
  const int grid = (int)(want > cap ? cap : want);
  adamw_flat_kernel<<<grid, OPT_THREADS, 0, (cudaStream_t)stream>>>(p, g, m, v, decay_bits, granules, hyper, gnorm_sq, loss,
                                                                   beta1, beta2, eps, weight_decay, skipped);
  return ogv_check_launch("adamw_flat");
}

extern "C" int ogv_train_metrics(const float* logits, long long ld, const long long* labels, int B, int K,
                                 const float* loss, float* acc, void* stream) {
  if (B == 0) return OGV_OK;
  OGV_REQUIRE(logits && labels && acc && K > 0 && ld >= K, "train_metrics: bad arguments");
  metrics_kernel<<<ogv_ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(logits, ld, labels, B, K, loss, acc);
  return ogv_check_launch("train_metrics");
}
