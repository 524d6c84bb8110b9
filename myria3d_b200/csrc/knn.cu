// Brute-force per-cloud k-NN for sm_100a.
//
// Replaces torch_cluster's knn (one thread per query scanning global memory)
// behind knn_graph(pos, k, batch, loop=True) (myria3d/models/modules/pyg_randla_net.py:180) and
// knn_interpolate (pyg_randla_net.py:250, myria3d/models/model.py:90).
//
// Design: a CTA owns 128 queries of ONE cloud; the cloud's candidate coordinates stream through a
// double-buffered shared-memory ring in 1024-point (12 KB) tiles staged by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx; the ragged last tile of the array falls back to plain
// loads); every lane reads the same candidate (LDS.128 broadcast, 4 candidates per 3 loads) and
// keeps its own sorted top-k in registers.  Distances use the reference's rounding sequence and
// candidates are visited in ascending index with a strict '<' insert, so ties resolve to the
// lower index and index sets are bit-exact against the oracle.
#include <math_constants.h>

#include "common.cuh"

namespace b200 {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;  // candidate points per shared-memory tile

template <int KMAX>
struct TopK {
  float d[KMAX];
  int idx[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      d[j] = CUDART_INF_F;
      idx[j] = -1;
    }
  }
  // insert (dist, i) keeping (d, idx) ascending; equal distances keep the earlier (lower) index first
  __device__ __forceinline__ void push(float dist, int i) {
    if (dist < d[KMAX - 1]) {
#pragma unroll
      for (int j = KMAX - 1; j >= 1; --j) {
        if (d[j - 1] > dist) {
          d[j] = d[j - 1];
          idx[j] = idx[j - 1];
        } else if (d[j] > dist) {
          d[j] = dist;
          idx[j] = i;
        }
      }
      if (d[0] > dist) {
        d[0] = dist;
        idx[0] = i;
      }
    }
  }
};

template <int KMAX>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float* __restrict__ pos_x, const int64_t* __restrict__ ptr_x, int64_t nx,
           const float* __restrict__ pos_y, const int64_t* __restrict__ ptr_y,
           int k, int kt, int32_t* __restrict__ nbr, float* __restrict__ dist2) {
  __shared__ __align__(16) float tile[2][KNN_TILE * 3];
  __shared__ __align__(8) uint64_t bars[2];

  const int cloud = blockIdx.y;
  const int64_t xs = ptr_x[cloud], xe = ptr_x[cloud + 1];
  const int64_t ys = ptr_y[cloud], ye = ptr_y[cloud + 1];
  const int64_t q0 = ys + (int64_t)blockIdx.x * KNN_THREADS;
  if (q0 >= ye) return;  // uniform for the whole CTA

  const int tid = threadIdx.x;
  const int64_t q = q0 + tid;
  const bool active = q < ye;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    qx = pos_y[3 * q + 0];
    qy = pos_y[3 * q + 1];
    qz = pos_y[3 * q + 2];
  }
  TopK<KMAX> top;
  top.init();

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // tiles are anchored at the cloud start rounded down to 4 points so that every bulk copy
  // starts on a 16-byte boundary (3 floats * 4 points = 48 B)
  const int64_t base = xs & ~int64_t(3);
  const int ntiles = (xe > base) ? (int)((xe - base + KNN_TILE - 1) / KNN_TILE) : 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(pos_x) & 15) == 0;

  auto is_bulk = [&](int t) -> bool {
    const int64_t start = base + (int64_t)t * KNN_TILE;
    return aligned && (start + KNN_TILE <= nx);
  };
  auto stage = [&](int t) {
    const int64_t start = base + (int64_t)t * KNN_TILE;
    float* dst = tile[t & 1];
    if (is_bulk(t)) {
      if (tid == 0) {
        mbar_expect_tx(&bars[t & 1], KNN_TILE * 12);
        tma_bulk_g2s(dst, pos_x + 3 * start, KNN_TILE * 12, &bars[t & 1]);
      }
    } else {
      const int64_t lim = (nx < start + KNN_TILE) ? nx : (start + KNN_TILE);
      const int cnt = (int)(lim - start) * 3;
      const float* src = pos_x + 3 * start;
      for (int i = tid; i < cnt; i += KNN_THREADS) dst[i] = src[i];
    }
  };

  uint32_t phase_bits = 0;
  if (ntiles > 0) {
    stage(0);
    if (!is_bulk(0)) __syncthreads();
  }
  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) stage(t + 1);
    if (is_bulk(t)) {
      mbar_wait(&bars[t & 1], (phase_bits >> (t & 1)) & 1u);
      phase_bits ^= (1u << (t & 1));
    }
    const int64_t start = base + (int64_t)t * KNN_TILE;
    const int jb = (xs > start) ? (int)(xs - start) : 0;
    const int je = (xe - start < KNN_TILE) ? (int)(xe - start) : KNN_TILE;
    if (active) {
      const float4* t4 = reinterpret_cast<const float4*>(tile[t & 1]);
      const int gend = (je + 3) >> 2;
      for (int g = jb >> 2; g < gend; ++g) {
        const float4 a = t4[3 * g + 0], b = t4[3 * g + 1], c = t4[3 * g + 2];
        const int j0 = g << 2;
        const int gi = (int)(start + j0);
        const float d0 = dist2_rn(a.x, a.y, a.z, qx, qy, qz);
        const float d1 = dist2_rn(a.w, b.x, b.y, qx, qy, qz);
        const float d2 = dist2_rn(b.z, b.w, c.x, qx, qy, qz);
        const float d3 = dist2_rn(c.y, c.z, c.w, qx, qy, qz);
        if (j0 >= jb && j0 + 4 <= je) {
          top.push(d0, gi + 0);
          top.push(d1, gi + 1);
          top.push(d2, gi + 2);
          top.push(d3, gi + 3);
        } else {
          if (j0 + 0 >= jb && j0 + 0 < je) top.push(d0, gi + 0);
          if (j0 + 1 >= jb && j0 + 1 < je) top.push(d1, gi + 1);
          if (j0 + 2 >= jb && j0 + 2 < je) top.push(d2, gi + 2);
          if (j0 + 3 >= jb && j0 + 3 < je) top.push(d3, gi + 3);
        }
      }
    }
    __syncthreads();  // everyone is done with buffer t&1 before it is refilled
  }

  if (active) {
    int32_t* orow = nbr + q * kt;
    float* drow = dist2 ? dist2 + q * kt : nullptr;
#pragma unroll
    for (int e = 0; e < KMAX; ++e) {
      if (e < kt) {
        const bool keep = (e < k) && (top.idx[e] >= 0);
        orow[e] = keep ? top.idx[e] : -1;
        if (drow) drow[e] = keep ? top.d[e] : CUDART_INF_F;
      }
    }
    for (int e = KMAX; e < kt; ++e) {
      orow[e] = -1;
      if (drow) drow[e] = CUDART_INF_F;
    }
  }
}

template <int KMAX>
static int launch_knn(const float* pos_x, const int64_t* ptr_x, int64_t nx, const float* pos_y,
                      const int64_t* ptr_y, int32_t num_clouds, int64_t max_q, int k, int kt, int32_t* nbr,
                      float* dist2, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(max_q, KNN_THREADS), (unsigned)num_clouds);
  knn_kernel<KMAX><<<grid, KNN_THREADS, 0, st>>>(pos_x, ptr_x, nx, pos_y, ptr_y, k, kt, nbr, dist2);
  B200_CHECK_LAUNCH("knn_kernel");
  return B200_OK;
}

}  // namespace b200

extern "C" int b200_knn(const float* pos_x, const int64_t* ptr_x, int64_t nx, const float* pos_y,
                        const int64_t* ptr_y, int64_t ny, int32_t num_clouds, int64_t max_queries_per_cloud,
                        int32_t k, int32_t kt, int32_t* nbr, float* dist2, void* stream) {
  using namespace b200;
  B200_REQUIRE(pos_x && ptr_x && pos_y && ptr_y && nbr, B200_E_INVALID, "b200_knn: null pointer");
  B200_REQUIRE(k >= 1 && kt >= k, B200_E_INVALID, "b200_knn: need 1 <= k <= kt (k=%d kt=%d)", k, kt);
  B200_REQUIRE(k <= 64, B200_E_UNSUPPORTED, "b200_knn: k=%d > 64 not supported", k);
  B200_REQUIRE(nx < (int64_t(1) << 31) && ny < (int64_t(1) << 31), B200_E_UNSUPPORTED, "b200_knn: more than 2^31 points");
  B200_REQUIRE(num_clouds >= 0 && num_clouds <= 65535, B200_E_UNSUPPORTED, "b200_knn: num_clouds=%d out of [0,65535]", num_clouds);
  if (ny == 0 || num_clouds == 0 || max_queries_per_cloud <= 0) return B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B200_KNN_CASE(KM) \
  if (k <= KM) return launch_knn<KM>(pos_x, ptr_x, nx, pos_y, ptr_y, num_clouds, max_queries_per_cloud, k, kt, nbr, dist2, st)
  B200_KNN_CASE(1);
  B200_KNN_CASE(2);
  B200_KNN_CASE(4);
  B200_KNN_CASE(8);
  B200_KNN_CASE(16);
  B200_KNN_CASE(32);
  B200_KNN_CASE(64);
#undef B200_KNN_CASE
  return B200_E_UNSUPPORTED;
}
