// Index / scatter kernels of the hot path for sm_100a (HBM-bound row moves, vectorised 16 B accesses).
//
//   gather_rows / scatter_rows_add  <- decimate(): tensor[idx_decim] and its backward
//                                      (myria3d/models/modules/pyg_randla_net.py:234-238)
//   knn_interp fwd / bwd            <- torch_geometric.nn.unpool.knn_interpolate's weighting tail
//                                      (pyg_randla_net.py:250, myria3d/models/model.py:90-98)
// Arithmetic follows the reference's rounding sequence so the forward results are bit-exact
// against the CPU oracle: w = 1/max(d2, 1e-16) (IEEE division), products x*w, sequential sums in
// ascending-distance order, one IEEE division by the weight sum.
#include "common.cuh"

namespace b200 {

template <int V>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ out,
                   int64_t n_out, int c) {
  const int cv = c / V;
  const int64_t total = n_out * cv;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int64_t r = t / cv;
    const int q = (int)(t % cv);
    const int64_t s = __ldg(idx + r);
    if constexpr (V == 4)
      reinterpret_cast<float4*>(out)[r * cv + q] = __ldg(reinterpret_cast<const float4*>(src) + s * cv + q);
    else
      out[r * cv + q] = __ldg(src + s * cv + q);
  }
}

template <int V>
__global__ void __launch_bounds__(256)
scatter_rows_add_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst,
                        int64_t n_src, int c) {
  const int cv = c / V;
  const int64_t total = n_src * cv;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int64_t r = t / cv;
    const int q = (int)(t % cv);
    const int64_t d = __ldg(idx + r);
    if constexpr (V == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + r * cv + q);
      atomicAdd(reinterpret_cast<float4*>(dst) + d * cv + q, v);
    } else {
      atomicAdd(dst + d * cv + q, __ldg(src + r * cv + q));
    }
  }
}

// y[i, 0:c) = (sum_e x[nbr_e] * w_e) / (sum_e w_e)
template <int V>
__global__ void __launch_bounds__(256)
knn_interp_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ nbr, const float* __restrict__ dist2,
                      float* __restrict__ out, int64_t ny, int c, int k, int kt, int64_t ld_out) {
  const int cv = c / V;
  const int64_t total = ny * cv;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int64_t i = t / cv;
    const int q = (int)(t % cv);
    float acc[V];
#pragma unroll
    for (int u = 0; u < V; ++u) acc[u] = 0.f;
    float wsum = 0.f;
    bool first = true;
    for (int e = 0; e < k; ++e) {
      const int j = __ldg(nbr + i * kt + e);
      if (j < 0) break;
      const float w = __fdiv_rn(1.0f, fmaxf(__ldg(dist2 + i * kt + e), 1e-16f));
      float xv[V];
      if constexpr (V == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)j * c) + q);
        xv[0] = v.x, xv[1] = v.y, xv[2] = v.z, xv[3] = v.w;
      } else {
        xv[0] = __ldg(x + (int64_t)j * c + q);
      }
#pragma unroll
      for (int u = 0; u < V; ++u) {
        const float p = __fmul_rn(xv[u], w);
        acc[u] = first ? p : __fadd_rn(acc[u], p);
      }
      wsum = first ? w : __fadd_rn(wsum, w);
      first = false;
    }
    float* o = out + i * ld_out + (int64_t)q * V;
#pragma unroll
    for (int u = 0; u < V; ++u) acc[u] = first ? 0.f : __fdiv_rn(acc[u], wsum);
    if constexpr (V == 4)
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
      o[0] = acc[0];
  }
}

// grad_x[nbr_e] += grad_y[i] * w_e / wsum   (autograd of (x*w summed) / wsum)
template <int V>
__global__ void __launch_bounds__(256)
knn_interp_bwd_kernel(const float* __restrict__ gy, int64_t ld_grad, const int32_t* __restrict__ nbr,
                      const float* __restrict__ dist2, float* __restrict__ gx, int64_t ny, int c, int k, int kt) {
  const int cv = c / V;
  const int64_t total = ny * cv;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int64_t i = t / cv;
    const int q = (int)(t % cv);
    float wsum = 0.f;
    bool first = true;
    for (int e = 0; e < k; ++e) {
      if (__ldg(nbr + i * kt + e) < 0) break;
      const float w = __fdiv_rn(1.0f, fmaxf(__ldg(dist2 + i * kt + e), 1e-16f));
      wsum = first ? w : __fadd_rn(wsum, w);
      first = false;
    }
    if (first) continue;
    float g[V];
    const float* gp = gy + i * ld_grad + (int64_t)q * V;
    if constexpr (V == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(gp));
      g[0] = v.x, g[1] = v.y, g[2] = v.z, g[3] = v.w;
    } else {
      g[0] = __ldg(gp);
    }
    for (int e = 0; e < k; ++e) {
      const int j = __ldg(nbr + i * kt + e);
      if (j < 0) break;
      const float w = __fdiv_rn(1.0f, fmaxf(__ldg(dist2 + i * kt + e), 1e-16f));
      float r[V];
#pragma unroll
      for (int u = 0; u < V; ++u) r[u] = __fmul_rn(__fdiv_rn(g[u], wsum), w);
      float* d = gx + (int64_t)j * c + (int64_t)q * V;
      if constexpr (V == 4)
        atomicAdd(reinterpret_cast<float4*>(d), make_float4(r[0], r[1], r[2], r[3]));
      else
        atomicAdd(d, r[0]);
    }
  }
}

static int grid_for(int64_t items) {
  int64_t blocks = ceil_div(items, 256);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b200

using namespace b200;

extern "C" int b200_gather_rows(const float* src, const int64_t* idx, float* out, int64_t n_out, int32_t c,
                                void* stream) {
  B200_REQUIRE(src && idx && out && c > 0, B200_E_INVALID, "b200_gather_rows: null pointer / c <= 0");
  if (n_out <= 0) return B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c % 4 == 0 && al16(src) && al16(out))
    gather_rows_kernel<4><<<grid_for(n_out * (c / 4)), 256, 0, st>>>(src, idx, out, n_out, c);
  else
    gather_rows_kernel<1><<<grid_for(n_out * c), 256, 0, st>>>(src, idx, out, n_out, c);
  B200_CHECK_LAUNCH("gather_rows_kernel");
  return B200_OK;
}

extern "C" int b200_scatter_rows_add(const float* src, const int64_t* idx, float* dst, int64_t n_src, int32_t c,
                                     void* stream) {
  B200_REQUIRE(src && idx && dst && c > 0, B200_E_INVALID, "b200_scatter_rows_add: null pointer / c <= 0");
  if (n_src <= 0) return B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c % 4 == 0 && al16(src) && al16(dst))
    scatter_rows_add_kernel<4><<<grid_for(n_src * (c / 4)), 256, 0, st>>>(src, idx, dst, n_src, c);
  else
    scatter_rows_add_kernel<1><<<grid_for(n_src * c), 256, 0, st>>>(src, idx, dst, n_src, c);
  B200_CHECK_LAUNCH("scatter_rows_add_kernel");
  return B200_OK;
}

extern "C" int b200_knn_interp_fwd(const float* x, const int32_t* nbr, const float* dist2, float* out, int64_t ny,
                                   int32_t c, int32_t k, int32_t kt, int64_t ld_out, void* stream) {
  B200_REQUIRE(x && nbr && dist2 && out && c > 0, B200_E_INVALID, "b200_knn_interp_fwd: null pointer / c <= 0");
  B200_REQUIRE(k >= 1 && kt >= k && ld_out >= c, B200_E_INVALID, "b200_knn_interp_fwd: need 1 <= k <= kt, ld_out >= c");
  if (ny <= 0) return B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c % 4 == 0 && ld_out % 4 == 0 && al16(x) && al16(out))
    knn_interp_fwd_kernel<4><<<grid_for(ny * (c / 4)), 256, 0, st>>>(x, nbr, dist2, out, ny, c, k, kt, ld_out);
  else
    knn_interp_fwd_kernel<1><<<grid_for(ny * c), 256, 0, st>>>(x, nbr, dist2, out, ny, c, k, kt, ld_out);
  B200_CHECK_LAUNCH("knn_interp_fwd_kernel");
  return B200_OK;
}

extern "C" int b200_knn_interp_bwd(const float* grad_y, int64_t ld_grad, const int32_t* nbr, const float* dist2,
                                   float* grad_x, int64_t ny, int32_t c, int32_t k, int32_t kt, void* stream) {
  B200_REQUIRE(grad_y && nbr && dist2 && grad_x && c > 0, B200_E_INVALID, "b200_knn_interp_bwd: null pointer / c <= 0");
  B200_REQUIRE(k >= 1 && kt >= k && ld_grad >= c, B200_E_INVALID, "b200_knn_interp_bwd: need 1 <= k <= kt, ld_grad >= c");
  if (ny <= 0) return B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c % 4 == 0 && ld_grad % 4 == 0 && al16(grad_y) && al16(grad_x))
    knn_interp_bwd_kernel<4><<<grid_for(ny * (c / 4)), 256, 0, st>>>(grad_y, ld_grad, nbr, dist2, grad_x, ny, c, k, kt);
  else
    knn_interp_bwd_kernel<1><<<grid_for(ny * c), 256, 0, st>>>(grad_y, ld_grad, nbr, dist2, grad_x, ny, c, k, kt);
  B200_CHECK_LAUNCH("knn_interp_bwd_kernel");
  return B200_OK;
}
