// Self-test of the tcgen05 / TMEM building blocks (tc.cuh): D[128, n] (+)= A[128, k] * B[n, k]^T with the
// operands split into tf32 hi/lo parts in shared memory and 1 (plain TF32) or 3 (3xTF32) accumulation passes
// into one TMEM accumulator.  Not on the hot path: it exists so that the descriptor / layout / tcgen05.ld /
// tcgen05.st conventions used by the fused kernels are pinned by a test against an fp64 product.
//
// passes = 1 / 3: kind::tf32 (plain / 3xTF32);  passes = 6: kind::f16 with bf16 x 3 operands (six cross products),
// the arithmetic of the fused LFA kernels.
// flags:  bit 0  A is staged "transposed" (buffer rows = k, 16-byte vectors along the 128 rows of A) and read
//                through the MN-major descriptor -- the way lfa_tc.cu re-reads dA / W_att without moving them
//                (bf16 only: a no-swizzle MN-major tf32 operand is not readable, see tc.cuh);
//         bit 1  the same for B;
//         bit 2  the accumulator is pre-initialised from d's incoming contents with tcgen05.st and every MMA
//                accumulates (the "direct term" of the fused LFA backward).
#include "tc.cuh"

namespace b200 {

__global__ void __launch_bounds__(128)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int n, int k,
                        int passes, int flags, uint32_t tmem_cols, int* __restrict__ status) {
  extern __shared__ __align__(128) float tc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const bool init_d = flags & 4;
  float* Ah = tc_smem;
  float* Al = Ah + tc::operand_floats(128, k);
  float* Bh = Al + tc::operand_floats(128, k);
  float* Bl = Bh + tc::operand_floats(n, k);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int t = tid; t < 128 * k; t += 128) {
    const int r = t % 128, kk = t / 128;
    float hi, lo;
    tc::split_tf32(A[r * k + kk], hi, lo);
    const int off = tc::operand_offset(128, r, kk);
    Ah[off] = hi;
    Al[off] = lo;
  }
  for (int t = tid; t < n * k; t += 128) {
    const int r = t % n, kk = t / n;
    float hi, lo;
    tc::split_tf32(B[r * k + kk], hi, lo);
    const int off = tc::operand_offset(n, r, kk);
    Bh[off] = hi;
    Bl[off] = lo;
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

  if (init_d) {  // accumulator <- D (thread = TMEM lane = row)
    for (int c0 = 0; c0 < n; c0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = D[tid * n + c0 + i];
      tc::tmem_st16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, n);
    const uint32_t lbo_a = tc::lbo_bytes(128), lbo_b = tc::lbo_bytes(n);
    bool acc = init_d;
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

__global__ void __launch_bounds__(128)
tc_gemm_selftest_bf16_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int n, int k,
                             int flags, uint32_t tmem_cols, int* __restrict__ status) {
  extern __shared__ __align__(128) uint16_t tcs16[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const bool a_mn = flags & 1, b_mn = flags & 2, init_d = flags & 4;
  // K-major plane: rows = MN index, chunked dimension = k.  MN-major plane: rows = k, chunked dimension = MN index.
  const int a_rows = a_mn ? k : 128, a_cdim = a_mn ? 128 : k;
  const int b_rows = b_mn ? k : n, b_cdim = b_mn ? n : k;
  const size_t a_plane = tc::plane_halves(a_rows, a_cdim), b_plane = tc::plane_halves(b_rows, b_cdim);
  uint16_t* Ap = tcs16;                // 3 planes
  uint16_t* Bp = tcs16 + 3 * a_plane;  // 3 planes
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int t = tid; t < 128 * k; t += 128) {
    const int r = t % 128, kk = t / 128;
    uint32_t t1, t2, t3;
    tc::split_bf16x3(A[r * k + kk], t1, t2, t3);
    const int off = a_mn ? tc::plane_offset(k, kk, r) : tc::plane_offset(128, r, kk);
    Ap[off] = (uint16_t)(t1 >> 16), Ap[a_plane + off] = (uint16_t)(t2 >> 16), Ap[2 * a_plane + off] = (uint16_t)(t3 >> 16);
  }
  for (int t = tid; t < n * k; t += 128) {
    const int r = t % n, kk = t / n;
    uint32_t t1, t2, t3;
    tc::split_bf16x3(B[r * k + kk], t1, t2, t3);
    const int off = b_mn ? tc::plane_offset(k, kk, r) : tc::plane_offset(n, r, kk);
    Bp[off] = (uint16_t)(t1 >> 16), Bp[b_plane + off] = (uint16_t)(t2 >> 16), Bp[2 * b_plane + off] = (uint16_t)(t3 >> 16);
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

  if (init_d) {
    for (int c0 = 0; c0 < n; c0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = D[tid * n + c0 + i];
      tc::tmem_st16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = tc::idesc_bf16(128, n, a_mn, b_mn);
    bool acc = init_d;
    for (int pass = 0; pass < 6; ++pass) {
      const uint32_t a_base = smem_u32(Ap + tc::bf16x3_term_a(pass) * a_plane);
      const uint32_t b_base = smem_u32(Bp + tc::bf16x3_term_b(pass) * b_plane);
      for (int k0 = 0; k0 < k; k0 += 16) {
        const uint64_t ad = a_mn ? tc::plane_desc_mn(a_base, a_rows, k0) : tc::plane_desc_k(a_base, a_rows, k0);
        const uint64_t bd = b_mn ? tc::plane_desc_mn(b_base, b_rows, k0) : tc::plane_desc_k(b_base, b_rows, k0);
        tc::mma_bf16(tmem_d, ad, bd, idesc, acc);
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
    for (int c0 = 0; c0 < n; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[tid * n + c0 + i] = v[i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_tc_gemm_selftest(const float* a, const float* b, float* d, int32_t n, int32_t k, int32_t passes,
                                     int32_t flags, int32_t* status, void* stream) {
  B200_REQUIRE(a && b && d && status, B200_E_INVALID, "b200_tc_gemm_selftest: null pointer");
  B200_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 8 && k % 8 == 0 && (passes == 1 || passes == 3 || passes == 6) &&
                   flags >= 0 && flags < 8,
               B200_E_INVALID,
               "b200_tc_gemm_selftest: need 16 <= n <= 256 (multiple of 16), k multiple of 8, passes in {1, 3, 6}, flags < 8");
  uint32_t cols = 32;
  while ((int)cols < n) cols <<= 1;
  if (passes == 6) {
    B200_REQUIRE(k % 16 == 0, B200_E_INVALID, "b200_tc_gemm_selftest: bf16 operands need k %% 16 == 0");
    const size_t a_plane = (flags & 1) ? tc::plane_halves(k, 128) : tc::plane_halves(128, k);
    const size_t b_plane = (flags & 2) ? tc::plane_halves(k, n) : tc::plane_halves(n, k);
    // the M = 128 / N = n reads of a K-major plane with a partial last row group stay inside the plane; MN-major
    // planes are read exactly
    const size_t smem = sizeof(uint16_t) * 3 * (a_plane + b_plane);
    B200_REQUIRE(smem <= 200 * 1024, B200_E_UNSUPPORTED, "b200_tc_gemm_selftest: operands need %zu bytes of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "tc selftest smem attribute");
    tc_gemm_selftest_bf16_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d, n, k, flags, cols, status);
    B200_CHECK_LAUNCH("tc_gemm_selftest_bf16_kernel");
    return B200_OK;
  }
  B200_REQUIRE((flags & 3) == 0, B200_E_UNSUPPORTED,
               "b200_tc_gemm_selftest: tf32 operands cannot be read MN-major from the no-swizzle layout (use passes = 6)");
  const size_t smem = sizeof(float) * 2 * (tc::operand_floats(128, k) + tc::operand_floats(n, k));
  B200_REQUIRE(smem <= 200 * 1024, B200_E_UNSUPPORTED, "b200_tc_gemm_selftest: operands need %zu bytes of shared memory", smem);
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "tc selftest smem attribute");
  tc_gemm_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d, n, k, passes, flags, cols, status);
  B200_CHECK_LAUNCH("tc_gemm_selftest_kernel");
  return B200_OK;
}
