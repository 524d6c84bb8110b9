// Self-test of the tcgen05 / TMEM building blocks (tc.cuh): D[128, n] = A[128, k] * B[n, k]^T with the
// operands split into tf32 hi/lo parts in shared memory and 1 (plain TF32) or 3 (3xTF32) accumulation passes
// into one TMEM accumulator.  Not on the hot path: it exists so that the descriptor / layout / tcgen05.ld
// conventions used by the fused kernels are pinned by a test against an fp64 product.
#include "tc.cuh"

namespace b200 {

__global__ void __launch_bounds__(128)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int n, int k,
                        int passes, uint32_t tmem_cols, int* __restrict__ status) {
  extern __shared__ __align__(128) float tc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* Ah = tc_smem;
  float* Al = Ah + tc::operand_floats(128, k);
  float* Bh = Al + tc::operand_floats(128, k);
  float* Bl = Bh + tc::operand_floats(n, k);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int t = tid; t < 128 * k; t += 128) {
    const int r = t % 128, kk = t / 128;
    float hi, lo;
    tc::split_tf32(A[r * k + kk], hi, lo);
    Ah[tc::operand_offset(128, r, kk)] = hi;
    Al[tc::operand_offset(128, r, kk)] = lo;
  }
  for (int t = tid; t < n * k; t += 128) {
    const int r = t % n, kk = t / n;
    float hi, lo;
    tc::split_tf32(B[r * k + kk], hi, lo);
    Bh[tc::operand_offset(n, r, kk)] = hi;
    Bl[tc::operand_offset(n, r, kk)] = lo;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, n);
    const uint32_t lbo_a = tc::lbo_bytes(128), lbo_b = tc::lbo_bytes(n);
    bool acc = false;
    for (int pass = 0; pass < passes; ++pass) {
      const float* a = (pass == 1) ? Al : Ah;  // hi*hi, lo*hi, hi*lo
      const float* b = (pass == 2) ? Bl : Bh;
      for (int k0 = 0; k0 < k; k0 += 8) {
        const uint64_t ad = tc::smem_desc(smem_u32(a) + (uint32_t)(k0 / 4) * lbo_a, lbo_a, tc::kSboBytes);
        const uint64_t bd = tc::smem_desc(smem_u32(b) + (uint32_t)(k0 / 4) * lbo_b, lbo_b, tc::kSboBytes);
        tc::mma_tf32(tmem_d, ad, bd, idesc, acc);
        acc = true;
      }
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait_bounded(&bar, 0);
  tc::fence_after_sync();
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    const int row = tid;  // TMEM lane == accumulator row
    for (int c0 = 0; c0 < n; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[row * n + c0 + i] = v[i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_tc_gemm_selftest(const float* a, const float* b, float* d, int32_t n, int32_t k, int32_t passes,
                                     int32_t* status, void* stream) {
  B200_REQUIRE(a && b && d && status, B200_E_INVALID, "b200_tc_gemm_selftest: null pointer");
  B200_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 8 && k % 8 == 0 && (passes == 1 || passes == 3), B200_E_INVALID,
               "b200_tc_gemm_selftest: need 16 <= n <= 256 (multiple of 16), k multiple of 8, passes in {1, 3}");
  const size_t smem = sizeof(float) * (2 * tc::operand_floats(128, k) + 2 * tc::operand_floats(n, k));
  B200_REQUIRE(smem <= 200 * 1024, B200_E_UNSUPPORTED, "b200_tc_gemm_selftest: operands need %zu bytes of shared memory", smem);
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "tc selftest smem attribute");
  uint32_t cols = 32;
  while ((int)cols < n) cols <<= 1;
  tc_gemm_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d, n, k, passes, cols, status);
  B200_CHECK_LAUNCH("tc_gemm_selftest_kernel");
  return B200_OK;
}
