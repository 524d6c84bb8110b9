// Flat Adam update for sm_100a: ONE kernel over the concatenated parameters (1.1 M floats) instead of torch's
// multi-tensor launches over 152 small tensors (the optimizer is not part of the reference's hot path --
// configs/model/optimizer/Adam.yaml just names torch.optim.Adam -- but it sits inside every measured step).
// Same arithmetic as torch.optim.Adam(betas, eps, weight_decay=0, amsgrad=False, maximize=False):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step counter lives on the device (int64) so that the update is CUDA-graph capturable.
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr_host, const float* __restrict__ lr_dev, float b1, float b2, float eps,
                 const int64_t* __restrict__ step) {
  const float lr = lr_dev ? *lr_dev : lr_host;  // device scalar: a captured graph follows the lr scheduler
  const double t = (double)(*step);
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 gq = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mq = reinterpret_cast<float4*>(m)[i], vq = reinterpret_cast<float4*>(v)[i], pq = reinterpret_cast<float4*>(p)[i];
    float gg[4] = {gq.x, gq.y, gq.z, gq.w}, mm[4] = {mq.x, mq.y, mq.z, mq.w}, vv[4] = {vq.x, vq.y, vq.z, vq.w},
          pp[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = b1 * mm[j] + (1.f - b1) * gg[j];
      vv[j] = b2 * vv[j] + (1.f - b2) * gg[j] * gg[j];
      pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi, v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

__global__ void increment_kernel(int64_t* step) { *step += 1; }

}  // namespace b200

using namespace b200;

extern "C" int b200_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              const float* lr_dev, float beta1, float beta2, float eps, int64_t* step, void* stream) {
  B200_REQUIRE(params && grads && exp_avg && exp_avg_sq && step, B200_E_INVALID, "b200_adam_flat: null pointer");
  B200_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, B200_E_INVALID,
               "b200_adam_flat: buffers must be 16-byte aligned");
  if (n <= 0) return B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  increment_kernel<<<1, 1, 0, st>>>(step);
  B200_CHECK_LAUNCH("increment_kernel");
  int64_t blocks = ceil_div(n / 4 + 1, 256);
  if (blocks > (int64_t)num_sms() * 4) blocks = (int64_t)num_sms() * 4;
  adam_flat_kernel<<<(unsigned)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, lr_dev, beta1, beta2, eps, step);
  B200_CHECK_LAUNCH("adam_flat_kernel");
  return B200_OK;
}
