// Random decimation draw: decimation_indices() of myria3d/models/modules/pyg_randla_net.py:192-231.
//
// The reference loops over the clouds of the batch in Python and calls torch.randperm(n_i)[:n_i // d] for each (two
// device->host syncs and ~7 launches per cloud and level).  Here ONE launch draws all clouds of a level: CTA b
// produces, for cloud b, a uniformly random ORDERED subset of k_b = new_ptr[b+1] - new_ptr[b] of its n_b points --
// the same distribution as randperm(n_b)[:k_b] -- without any host involvement, so the draw can live inside a
// captured CUDA graph (the random stream advances through a device-side counter, like adam_flat's step counter).
//
//   key(i) = Philox4x32-10(counter = (i, draw counter), key = (seed, level salt)).x     one 32-bit key per point
//   the k_b points with the smallest (key, index) pairs, in ascending (key, index) order, are the draw:
//     1. radix select (4 passes of 8-bit shared-memory histograms) finds the k_b-th smallest key; keys are
//        recomputed on the fly instead of stored (a 65 536-point cloud would not fit otherwise),
//     2. the selected (key, index) pairs are compacted into shared memory,
//     3. a bitonic sort orders them, 4. index + ptr[b] is written as int64.
// Equal keys (expected 2e-2 pairs per 12 800-point cloud) are ordered by index, so the result is a deterministic
// function of (seed, counter, salt, batch layout).
#include "common.cuh"

namespace b200 {

constexpr int DD_THREADS = 1024;

__device__ __forceinline__ uint32_t philox_key(uint32_t i_lo, uint32_t i_hi, uint32_t c_lo, uint32_t c_hi, uint32_t k0, uint32_t k1) {
  uint32_t x0 = i_lo, x1 = i_hi, x2 = c_lo, x3 = c_hi;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
    x0 = y0, x1 = y1, x2 = y2, x3 = y3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return x0;
}

__global__ void __launch_bounds__(DD_THREADS)
decimation_draw_kernel(const int64_t* __restrict__ ptr, const int64_t* __restrict__ new_ptr, int64_t* __restrict__ idx_out,
                       uint64_t seed, const int64_t* __restrict__ counter, uint32_t salt) {
  extern __shared__ __align__(16) unsigned long long dd_list[];  // [sort_capacity] (key << 32 | local index)
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh_bin, sh_below, sh_count;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t p0 = ptr[b];
  const int n = (int)(ptr[b + 1] - p0);
  const int64_t q0 = new_ptr[b];
  const int k = (int)(new_ptr[b + 1] - q0);
  if (n <= 0 || k <= 0) return;
  const uint64_t ctr = (uint64_t)(counter ? *counter : 0);
  const uint32_t c_lo = (uint32_t)ctr, c_hi = (uint32_t)(ctr >> 32);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ salt;
  auto key_of = [&](int i) {
    const uint64_t g = (uint64_t)(p0 + i);  // global point index: clouds never share keys
    return philox_key((uint32_t)g, (uint32_t)(g >> 32), c_lo, c_hi, k0, k1);
  };

  // ---- 1. radix select: the k-th smallest key T and how many keys equal to T belong to the draw
  uint32_t prefix = 0, mask = 0;
  unsigned int remaining = (unsigned int)k;  // rank of the wanted element among the keys matching the prefix
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += DD_THREADS) {
      const uint32_t key = key_of(i);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {  // one warp: inclusive scan over the 256 bins, 8 per lane
      unsigned int c[8], s = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        c[u] = hist[tid * 8 + u];
        s += c[u];
      }
      unsigned int incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      unsigned int below = incl - s;  // keys in bins before this lane's
      if (below < remaining && remaining <= incl) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (below < remaining && remaining <= below + c[u]) {
            sh_bin = (unsigned int)(tid * 8 + u);
            sh_below = below;
          }
          below += c[u];
        }
      }
    }
    __syncthreads();
    prefix |= sh_bin << shift;
    mask |= 255u << shift;
    remaining -= sh_below;
    __syncthreads();
  }
  const uint32_t T = prefix;
  const unsigned int take_eq = remaining;  // >= 1: keys equal to T that are part of the draw (lowest indices first)
  const unsigned int cnt_eq = hist[sh_bin];

  // ---- 2. compaction of the selected (key, index) pairs
  if (tid == 0) sh_count = 0;
  __syncthreads();
  for (int i = tid; i < n; i += DD_THREADS) {
    const uint32_t key = key_of(i);
    if (key < T || (key == T && cnt_eq == take_eq)) {
      const unsigned int slot = atomicAdd(&sh_count, 1u);
      dd_list[slot] = ((unsigned long long)key << 32) | (unsigned int)i;
    }
  }
  if (cnt_eq != take_eq && tid < 32) {  // (rare) more ties at the threshold than places: lowest indices win
    unsigned int taken = 0;
    for (int base = 0; base < n && taken < take_eq; base += 32) {
      const int i = base + tid;
      const bool eq = (i < n) && key_of(i) == T;
      const unsigned int m = __ballot_sync(0xffffffffu, eq);
      const unsigned int rank = taken + __popc(m & ((1u << tid) - 1u));
      if (eq && rank < take_eq) {
        const unsigned int slot = atomicAdd(&sh_count, 1u);
        dd_list[slot] = ((unsigned long long)T << 32) | (unsigned int)i;
      }
      taken += __popc(m);
    }
  }
  __syncthreads();
  // ---- 3. bitonic sort of the k pairs (padded to a power of two with +inf)
  int cap = 1;
  while (cap < k) cap <<= 1;
  for (int i = k + tid; i < cap; i += DD_THREADS) dd_list[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= cap; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (cap >> 1); t += DD_THREADS) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool ascending = ((lo & size) == 0);
        const unsigned long long a = dd_list[lo], c = dd_list[hi];
        if ((a > c) == ascending) {
          dd_list[lo] = c;
          dd_list[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  // ---- 4. the draw, as indices into the batch
  for (int t = tid; t < k; t += DD_THREADS) idx_out[q0 + t] = p0 + (int64_t)(unsigned int)(dd_list[t] & 0xffffffffull);
}

__global__ void counter_add_kernel(int64_t* counter, int64_t delta) { *counter += delta; }

}  // namespace b200

extern "C" int b200_decimation_draw(const int64_t* ptr, const int64_t* new_ptr, int32_t num_clouds, int64_t max_kept,
                                    uint64_t seed, const int64_t* counter, uint32_t salt, int64_t* idx_out, void* stream) {
  using namespace b200;
  B200_REQUIRE(ptr && new_ptr && idx_out, B200_E_INVALID, "b200_decimation_draw: null pointer");
  B200_REQUIRE(num_clouds >= 0 && max_kept >= 0, B200_E_INVALID, "b200_decimation_draw: negative size");
  if (num_clouds == 0 || max_kept == 0) return B200_OK;
  int64_t cap = 1;
  while (cap < max_kept) cap <<= 1;
  const size_t smem = (size_t)cap * sizeof(unsigned long long);
  B200_REQUIRE(smem <= 200 * 1024, B200_E_UNSUPPORTED,
               "b200_decimation_draw: %lld kept points per cloud need %zu bytes of shared memory (max 25600 points)",
               (long long)max_kept, smem);
  cudaError_t e = cudaFuncSetAttribute(decimation_draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "decimation_draw smem attribute");
  decimation_draw_kernel<<<(unsigned)num_clouds, DD_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      ptr, new_ptr, idx_out, seed, counter, salt);
  B200_CHECK_LAUNCH("decimation_draw_kernel");
  return B200_OK;
}

extern "C" int b200_counter_add(int64_t* counter, int64_t delta, void* stream) {
  using namespace b200;
  B200_REQUIRE(counter, B200_E_INVALID, "b200_counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(counter, delta);
  B200_CHECK_LAUNCH("counter_add_kernel");
  return B200_OK;
}
