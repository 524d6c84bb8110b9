// Sliding-window stitch of per-sample logits back onto the full cloud for sm_100a.
//
//   stitch_segment_sum  <- torch_scatter.scatter_sum(logits, idx_in_full_cloud, out=zeros(nb_points, C))
//                          (myria3d/models/interpolation.py:113-116)
//   stitch_finalize     <- reduced_logits[idx_in_full_cloud] (:121), Softmax(dim=1) (:142), argmax (:145),
//                          Categorical(probs=probas).entropy() (:166)
//
// The reference's CPU scatter adds the contributions of a point in INPUT order onto a zero row; fp32 addition is not
// associative, so with overlapping windows (up to 4 predictions per point at overlap 25 m) an atomics-based scatter
// would differ in the last bit from run to run.  Here the caller supplies a STABLE argsort of idx; each destination
// row is then summed sequentially in input order by one thread per (row, 16-byte column group): deterministic and
// bit-identical to the CPU result.  Both kernels are HBM-bound row moves.
#include <math_constants.h>

#include "common.cuh"

namespace b200 {

static inline int grid_for_items(int64_t items) {
  int64_t g = ceil_div(items, 256);
  int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int V>
__global__ void __launch_bounds__(256)
stitch_segment_sum_kernel(const float* __restrict__ rows, const int64_t* __restrict__ order,
                          const int64_t* __restrict__ sorted_idx, float* __restrict__ out, int64_t m, int c,
                          int64_t nb_points) {
  const int cv = c / V;
  const int64_t total = m * cv;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int64_t p = t / cv;
    const int q = (int)(t % cv);
    const int64_t dst = __ldg(sorted_idx + p);
    if (p > 0 && __ldg(sorted_idx + p - 1) == dst) continue;  // not a segment head
    if (dst < 0 || dst >= nb_points) continue;                 // the wrapper validates; never write out of bounds
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    for (int64_t e = p; e < m && __ldg(sorted_idx + e) == dst; ++e) {
      const int64_t r = __ldg(order + e);
      if constexpr (V == 4) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(rows) + r * cv + q);
        acc[0] = __fadd_rn(acc[0], x.x), acc[1] = __fadd_rn(acc[1], x.y);
        acc[2] = __fadd_rn(acc[2], x.z), acc[3] = __fadd_rn(acc[3], x.w);
      } else {
        acc[0] = __fadd_rn(acc[0], __ldg(rows + r * cv + q));
      }
    }
    if constexpr (V == 4)
      reinterpret_cast<float4*>(out)[dst * cv + q] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
      out[dst * cv + q] = acc[0];
  }
}

// one thread per output row; c <= 32 classes live in registers
constexpr int kMaxClasses = 32;

__global__ void __launch_bounds__(256)
stitch_finalize_kernel(const float* __restrict__ reduced, const int64_t* __restrict__ idx, float* __restrict__ logits,
                       float* __restrict__ probas, int64_t* __restrict__ preds, float* __restrict__ entropy, int64_t m,
                       int c) {
  constexpr float kEps = 1.1920928955078125e-07f;  // torch.finfo(torch.float32).eps (probs_to_logits clamp)
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < m; i += (int64_t)gridDim.x * 256) {
    const float* src = reduced + __ldg(idx + i) * c;
    float v[kMaxClasses];
    float mx = -CUDART_INF_F;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) {
        v[k] = __ldg(src + k);
        if (v[k] > mx) mx = v[k], arg = k;  // strict: first maximum, like torch.argmax on CPU
      }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) {
        if (logits) logits[i * c + k] = v[k];
        v[k] = expf(v[k] - mx);
        sum += v[k];
      }
    float psum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < c) {
        v[k] = v[k] / sum;
        psum += v[k];
        if (probas) probas[i * c + k] = v[k];
      }
    if (preds) preds[i] = arg;
    if (entropy) {
      // torch.distributions.Categorical(probs=p): p <- p / p.sum(-1); logits = log(clamp(p, eps, 1 - eps));
      // entropy = -sum(logits * p)
      float h = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < c) {
          const float p = v[k] / psum;
          h += logf(fminf(fmaxf(p, kEps), 1.f - kEps)) * p;
        }
      entropy[i] = -h;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_stitch_segment_sum(const float* rows, const int64_t* order, const int64_t* sorted_idx, float* out,
                                       int64_t m, int32_t c, int64_t nb_points, void* stream) {
  B200_REQUIRE(c > 0 && nb_points >= 0, B200_E_INVALID, "b200_stitch_segment_sum: c <= 0 or nb_points < 0");
  if (m <= 0) return B200_OK;
  B200_REQUIRE(rows && order && sorted_idx && out, B200_E_INVALID, "b200_stitch_segment_sum: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = c % 4 == 0 && ((uintptr_t)rows % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (vec)
    stitch_segment_sum_kernel<4><<<grid_for_items(m * (c / 4)), 256, 0, st>>>(rows, order, sorted_idx, out, m, c, nb_points);
  else
    stitch_segment_sum_kernel<1><<<grid_for_items(m * c), 256, 0, st>>>(rows, order, sorted_idx, out, m, c, nb_points);
  B200_CHECK_LAUNCH("stitch_segment_sum_kernel");
  return B200_OK;
}

extern "C" int b200_stitch_finalize(const float* reduced, const int64_t* idx, float* logits, float* probas, int64_t* preds,
                                    float* entropy, int64_t m, int32_t c, void* stream) {
  B200_REQUIRE(c > 0 && c <= kMaxClasses, B200_E_UNSUPPORTED, "b200_stitch_finalize: 1 <= c <= %d classes (got %d)",
               kMaxClasses, (int)c);
  if (m <= 0) return B200_OK;
  B200_REQUIRE(reduced && idx, B200_E_INVALID, "b200_stitch_finalize: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  stitch_finalize_kernel<<<grid_for_items(m), 256, 0, st>>>(reduced, idx, logits, probas, preds, entropy, m, c);
  B200_CHECK_LAUNCH("stitch_finalize_kernel");
  return B200_OK;
}
