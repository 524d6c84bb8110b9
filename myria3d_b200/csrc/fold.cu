// Encoder fold: Linear(10 -> h) + BatchNorm of LocalFeatureAggregation.mlp_encoder
// (myria3d/models/modules/pyg_randla_net.py:117,144) as ONE tiny kernel per direction.
//
// The reference feeds r = [p_i, p_j, p_j - p_i, dist] (:143); W r = Wq q with q = (p_i, p_j, dist) and
// Wq = [W_a - W_c, W_b + W_c, w_d].  Train-mode BatchNorm statistics over all E edges follow from the
// fp64 edge moments (b200_edge_moments): mean = Wq mu + b, var = diag(Wq C Wq^T).
// The fused LFA kernels then see the affine map  z = enc_w q + enc_b  with
//   enc_w = Wq * s,  enc_b = (b - mean) * s + beta,  s = gamma / sqrt(var + eps).
// The backward below is the exact gradient of that map w.r.t. W, b, gamma, beta (statistics included),
// i.e. what autograd computes through Linear + BatchNorm1d in training mode.
#include "common.cuh"

namespace b200 {

struct FoldRow {
  double wq[7];
  double mu[7];
  double cw[7];  // C wq
  double u, v;   // wq.mu, wq^T C wq
};

__device__ __forceinline__ void fold_row(const float* __restrict__ w, const double* __restrict__ mom, int m, FoldRow& R) {
  const float* wr = w + m * 10;
  R.wq[0] = (double)wr[0] - (double)wr[6];
  R.wq[1] = (double)wr[1] - (double)wr[7];
  R.wq[2] = (double)wr[2] - (double)wr[8];
  R.wq[3] = (double)wr[3] + (double)wr[6];
  R.wq[4] = (double)wr[4] + (double)wr[7];
  R.wq[5] = (double)wr[5] + (double)wr[8];
  R.wq[6] = (double)wr[9];
  R.u = 0.0, R.v = 0.0;
  if (mom) {
    const double inv_e = 1.0 / mom[0];
#pragma unroll
    for (int a = 0; a < 7; ++a) R.mu[a] = mom[1 + a] * inv_e;
#pragma unroll
    for (int a = 0; a < 7; ++a) {
      double acc = 0.0;
#pragma unroll
      for (int b = 0; b < 7; ++b) acc += (mom[8 + a * 7 + b] * inv_e - R.mu[a] * R.mu[b]) * R.wq[b];
      R.cw[a] = acc;
      R.u += R.wq[a] * R.mu[a];
      R.v += R.wq[a] * acc;
    }
    if (R.v < 0.0) R.v = 0.0;
  }
}

__global__ void encoder_fold_fwd_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const double* __restrict__ mom, float* __restrict__ running_mean,
                                        float* __restrict__ running_var, int64_t* __restrict__ num_batches_tracked,
                                        float momentum, float eps, float* __restrict__ enc_w,
                                        float* __restrict__ enc_b, int h) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m == 0 && mom && num_batches_tracked) *num_batches_tracked += 1;
  if (m >= h) return;
  FoldRow R;
  fold_row(w, mom, m, R);
  const double bias = b ? (double)b[m] : 0.0;
  double mean, var;
  if (mom) {
    mean = R.u + bias;
    var = R.v;
    if (running_mean) {
      const double e = mom[0];
      const double unbiased = (e > 1.0) ? var * e / (e - 1.0) : var;
      running_mean[m] = (float)((1.0 - (double)momentum) * (double)running_mean[m] + (double)momentum * mean);
      running_var[m] = (float)((1.0 - (double)momentum) * (double)running_var[m] + (double)momentum * unbiased);
    }
  } else {
    mean = (double)running_mean[m];
    var = (double)running_var[m];
  }
  const double s = (double)gamma[m] / sqrt(var + (double)eps);
#pragma unroll
  for (int t = 0; t < 7; ++t) enc_w[m * 7 + t] = (float)(R.wq[t] * s);
  enc_b[m] = (float)((bias - mean) * s + (double)beta[m]);
}

__global__ void encoder_fold_bwd_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                        const float* __restrict__ gamma, const double* __restrict__ mom,
                                        const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                        float eps, const float* __restrict__ g_enc_w, const float* __restrict__ g_enc_b,
                                        float* __restrict__ grad_w, float* __restrict__ grad_b,
                                        float* __restrict__ grad_gamma, float* __restrict__ grad_beta, int h,
                                        int accumulate) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= h) return;
  FoldRow R;
  fold_row(w, mom, m, R);
  const double bias = b ? (double)b[m] : 0.0;
  const double var = mom ? R.v : (double)running_var[m];
  const double mean = mom ? (R.u + bias) : (double)running_mean[m];
  const double r = 1.0 / sqrt(var + (double)eps);
  const double g = (double)gamma[m];
  const double s = g * r;
  const double gc = (double)g_enc_b[m];
  double ga[7];
  double ds = gc * (bias - mean);
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    ga[t] = (double)g_enc_w[m * 7 + t];
    ds += ga[t] * R.wq[t];
  }
  double dwq[7];
  const double dv = mom ? (-0.5 * ds * g * r * r * r) : 0.0;
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    dwq[t] = ga[t] * s;
    if (mom) dwq[t] += -gc * s * R.mu[t] + 2.0 * dv * R.cw[t];
  }
  float* gw = grad_w + m * 10;
  const float gwv[10] = {(float)dwq[0], (float)dwq[1], (float)dwq[2], (float)dwq[3], (float)dwq[4], (float)dwq[5],
                         (float)(dwq[3] - dwq[0]), (float)(dwq[4] - dwq[1]), (float)(dwq[5] - dwq[2]), (float)dwq[6]};
  const float gbv = mom ? 0.f : (float)(gc * s);  // train mode: the bias cancels inside BatchNorm
  if (accumulate) {  // outputs are the parameters' .grad buffers
#pragma unroll
    for (int j = 0; j < 10; ++j) gw[j] += gwv[j];
    if (grad_b) grad_b[m] += gbv;
    grad_gamma[m] += (float)(ds * r);
    grad_beta[m] += (float)gc;
  } else {
#pragma unroll
    for (int j = 0; j < 10; ++j) gw[j] = gwv[j];
    if (grad_b) grad_b[m] = gbv;
    grad_gamma[m] = (float)(ds * r);
    grad_beta[m] = (float)gc;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_encoder_fold_fwd(const float* w, const float* b, const float* gamma, const float* beta,
                                     const double* moments, float* running_mean, float* running_var,
                                     int64_t* num_batches_tracked, float momentum, float eps, float* enc_w,
                                     float* enc_b, int32_t h, void* stream) {
  B200_REQUIRE(w && gamma && beta && enc_w && enc_b && h > 0, B200_E_INVALID, "b200_encoder_fold_fwd: null pointer / h <= 0");
  B200_REQUIRE(moments || (running_mean && running_var), B200_E_INVALID,
               "b200_encoder_fold_fwd: eval mode needs running statistics");
  B200_REQUIRE((running_mean == nullptr) == (running_var == nullptr), B200_E_INVALID,
               "b200_encoder_fold_fwd: running_mean / running_var must come together");
  encoder_fold_fwd_kernel<<<(unsigned)ceil_div(h, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      w, b, gamma, beta, moments, running_mean, running_var, num_batches_tracked, momentum, eps, enc_w, enc_b, h);
  B200_CHECK_LAUNCH("encoder_fold_fwd_kernel");
  return B200_OK;
}

extern "C" int b200_encoder_fold_bwd(const float* w, const float* b, const float* gamma, const double* moments,
                                     const float* running_mean, const float* running_var, float eps,
                                     const float* g_enc_w, const float* g_enc_b, float* grad_w, float* grad_b,
                                     float* grad_gamma, float* grad_beta, int32_t h, int32_t accumulate,
                                     void* stream) {
  B200_REQUIRE(w && gamma && g_enc_w && g_enc_b && grad_w && grad_gamma && grad_beta && h > 0, B200_E_INVALID,
               "b200_encoder_fold_bwd: null pointer / h <= 0");
  B200_REQUIRE(moments || (running_mean && running_var), B200_E_INVALID,
               "b200_encoder_fold_bwd: eval mode needs running statistics");
  encoder_fold_bwd_kernel<<<(unsigned)ceil_div(h, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      w, b, gamma, moments, running_mean, running_var, eps, g_enc_w, g_enc_b, grad_w, grad_b, grad_gamma, grad_beta, h,
      accumulate);
  B200_CHECK_LAUNCH("encoder_fold_bwd_kernel");
  return B200_OK;
}
