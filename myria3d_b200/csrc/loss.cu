// Cross-entropy of the training/validation step for sm_100a:
//   torch.nn.CrossEntropyLoss(weight=w | None, ignore_index=65, label_smoothing=0, reduction="mean")(logits, targets)
//   (configs/model/criterion/*.yaml, myria3d/models/model.py:117-118,135-136,152-153)
//
//   loss = sum_i w[t_i] * (logsumexp(x_i) - x_i[t_i]) / sum_i w[t_i]          over rows with t_i != ignore_index
//   dL/dx_i = g * w[t_i] * (softmax(x_i) - onehot(t_i)) / sum_i w[t_i]
//
// torch's own nll_loss kernels reduce [N, C] with a single CTA (0.33 ms for N = 204 800 in the step profile); here one
// thread owns a row (C <= 32 classes in registers), CTAs reduce in fp64 and the last CTA to finish writes the mean.
#include <math_constants.h>

#include "common.cuh"

namespace b200 {

constexpr int kCeMaxClasses = 32;
constexpr int kCeThreads = 256;

// acc[0] = sum of weighted losses, acc[1] = sum of weights, counter = CTAs done  (all pre-zeroed by the caller)
__global__ void __launch_bounds__(kCeThreads)
cross_entropy_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                         const float* __restrict__ weight, int64_t n, int c, int64_t ignore_index,
                         double* __restrict__ acc, unsigned int* __restrict__ counter, float* __restrict__ loss_out) {
  double sl = 0.0, sw = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kCeThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCeThreads) {
    const int64_t t = __ldg(target + i);
    if (t == ignore_index) continue;
    if (t < 0 || t >= c) {  // torch raises a device-side assert here; poison the loss instead of reading out of bounds
      sl = CUDART_NAN;
      continue;
    }
    const float* row = logits + i * c;
    float v[kCeMaxClasses];
    float mx = -CUDART_INF_F, xt = 0.f;
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k)
      if (k < c) {
        v[k] = __ldg(row + k);
        mx = fmaxf(mx, v[k]);
        if (k == (int)t) xt = v[k];
      }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k)
      if (k < c) s += expf(v[k] - mx);
    const float w = weight ? __ldg(weight + t) : 1.f;
    sl += (double)(w * ((mx - xt) + logf(s)));  // -log_softmax(x)[t]
    sw += (double)w;
  }
  __shared__ double sh[2][kCeThreads / 32];
  sl = warp_sum(sl), sw = warp_sum(sw);
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = sl, sh[1][threadIdx.x >> 5] = sw;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kCeThreads / 32; ++w) a += sh[0][w], b += sh[1][w];
    atomicAdd(acc + 0, a);
    atomicAdd(acc + 1, b);
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {  // last CTA: every partial is visible
      __threadfence();
      const double tl = atomicAdd(acc + 0, 0.0), tw = atomicAdd(acc + 1, 0.0);
      loss_out[0] = (float)(tl / tw);  // 0/0 = NaN when every row is ignored, like torch
      loss_out[1] = (float)tw;         // kept for the backward pass
    }
  }
}

template <int V>
__global__ void __launch_bounds__(kCeThreads)
cross_entropy_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                         const float* __restrict__ weight, int64_t n, int c, int64_t ignore_index,
                         const float* __restrict__ loss_and_wsum, const float* __restrict__ grad_loss,
                         float* __restrict__ grad_logits) {
  const float scale = __ldg(grad_loss) / __ldg(loss_and_wsum + 1);
  for (int64_t i = (int64_t)blockIdx.x * kCeThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kCeThreads) {
    const int64_t t = __ldg(target + i);
    const float* row = logits + i * c;
    float* grow = grad_logits + i * c;
    const bool live = t != ignore_index && t >= 0 && t < c;
    float v[kCeMaxClasses];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k)
      if (k < c) {
        v[k] = live ? __ldg(row + k) : 0.f;
        mx = fmaxf(mx, v[k]);
      }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k)
      if (k < c) {
        v[k] = expf(v[k] - mx);
        s += v[k];
      }
    const float f = live ? scale * (weight ? __ldg(weight + t) : 1.f) : 0.f;
    const float inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k)
      if (k < c) grow[k] = f * (v[k] * inv - (k == (int)t ? 1.f : 0.f));
  }
}

static inline int ce_grid(int64_t n) {
  int64_t g = ceil_div(n, kCeThreads);
  const int64_t cap = (int64_t)num_sms() * 4;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace b200

using namespace b200;

extern "C" int b200_cross_entropy_fwd(const float* logits, const int64_t* target, const float* weight, int64_t n,
                                      int32_t c, int64_t ignore_index, double* acc, uint32_t* counter, float* loss_out,
                                      void* stream) {
  B200_REQUIRE(c > 0 && c <= kCeMaxClasses, B200_E_UNSUPPORTED, "b200_cross_entropy_fwd: 1 <= c <= %d classes (got %d)",
               kCeMaxClasses, (int)c);
  B200_REQUIRE(n >= 0 && acc && counter && loss_out && (n == 0 || (logits && target)), B200_E_INVALID,
               "b200_cross_entropy_fwd: null pointer / n < 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cross_entropy_fwd_kernel<<<ce_grid(n), kCeThreads, 0, st>>>(logits, target, weight, n, c, ignore_index, acc, counter,
                                                               loss_out);
  B200_CHECK_LAUNCH("cross_entropy_fwd_kernel");
  return B200_OK;
}

extern "C" int b200_cross_entropy_bwd(const float* logits, const int64_t* target, const float* weight, int64_t n,
                                      int32_t c, int64_t ignore_index, const float* loss_and_wsum, const float* grad_loss,
                                      float* grad_logits, void* stream) {
  B200_REQUIRE(c > 0 && c <= kCeMaxClasses, B200_E_UNSUPPORTED, "b200_cross_entropy_bwd: 1 <= c <= %d classes (got %d)",
               kCeMaxClasses, (int)c);
  if (n <= 0) return B200_OK;
  B200_REQUIRE(logits && target && loss_and_wsum && grad_loss && grad_logits, B200_E_INVALID,
               "b200_cross_entropy_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cross_entropy_bwd_kernel<1><<<ce_grid(n), kCeThreads, 0, st>>>(logits, target, weight, n, c, ignore_index, loss_and_wsum,
                                                                 grad_loss, grad_logits);
  B200_CHECK_LAUNCH("cross_entropy_bwd_kernel");
  return B200_OK;
}
