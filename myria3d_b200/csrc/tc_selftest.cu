// Self-test of the tcgen05 / TMEM building blocks (tc.cuh): D[128, n] (+)= A[128, k] * B[n, k]^T with the
// operands split into tf32 hi/lo parts in shared memory and 1 (plain TF32) or 3 (3xTF32) accumulation passes
// into one TMEM accumulator.  Not on the hot path: it exists so that the descriptor / layout / tcgen05.ld /
// tcgen05.st conventions used by the fused kernels are pinned by a test against an fp64 product.
//
// passes = 1 / 3: kind::tf32 (plain / 3xTF32);  passes = 6: kind::f16 with bf16 x 3 operands (six cross products, the
// arithmetic of tc_skinny.cu);  passes = 2: kind::f16 with fp16 x 2 operands (three cross products, the arithmetic of
// the fused LFA kernels; operands in fp16's normal range).
// flags:  bit 0  A is staged "transposed" (buffer rows = k, 16-byte vectors along the 128 rows of A) and read
//                through the MN-major descriptor -- the way lfa_tc.cu re-reads dA / W_att without moving them
//                (bf16 only: a no-swizzle MN-major tf32 operand is not readable, see tc.cuh);
//         bit 1  the same for B;
//         bit 2  the accumulator is pre-initialised from d's incoming contents with tcgen05.st and every MMA
//                accumulates (the "direct term" of the fused LFA backward);
//         bit 3  (passes = 6, k % 32 == 0, not with bit 0) A lives in TENSOR MEMORY, written with tcgen05.st -- how
//                lfa_tc.cu keeps W_att resident next to the accumulators.
#include "tc.cuh"

namespace b200 {
long long* tc_debug_buffer();  // runtime.cu

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

template <int KC, bool F16>  // KC > 0: k is the compile-time constant KC and the issue loop is unrolled (scripts/mma_probe.py)
__global__ void __launch_bounds__(128)                                  // F16: fp16 x 2 planes / 3 products instead of bf16 x 3 / 6
tc_gemm_selftest_bf16_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int n, int k,
                             int flags, uint32_t tmem_cols, int* __restrict__ status, long long* __restrict__ dbg) {
  extern __shared__ __align__(128) uint16_t tcs16[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const bool a_mn = flags & 1, b_mn = flags & 2, init_d = flags & 4, a_tmem = flags & 8;
  // K-major plane: rows = MN index, chunked dimension = k.  MN-major plane: rows = k, chunked dimension = MN index.
  const int a_rows = a_mn ? k : 128, a_cdim = a_mn ? 128 : k;
  const int b_rows = b_mn ? k : n, b_cdim = b_mn ? n : k;
  const size_t a_plane = tc::plane_halves(a_rows, a_cdim), b_plane = tc::plane_halves(b_rows, b_cdim);
  uint16_t* Ap = tcs16;                // 3 planes
  uint16_t* Bp = tcs16 + 3 * a_plane;  // 3 planes
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int t = tid; t < 128 * k; t += 128) {
    const int r = t % 128, kk = t / 128;
    const int off = a_mn ? tc::plane_offset(k, kk, r) : tc::plane_offset(128, r, kk);
    if constexpr (F16) {
      uint32_t hi, lo;
      tc::split_f16x2_pair(A[r * k + kk], 0.f, hi, lo);
      Ap[off] = (uint16_t)hi, Ap[a_plane + off] = (uint16_t)lo;
    } else {
      uint32_t t1, t2, t3;
      tc::split_bf16x3(A[r * k + kk], t1, t2, t3);
      Ap[off] = (uint16_t)(t1 >> 16), Ap[a_plane + off] = (uint16_t)(t2 >> 16), Ap[2 * a_plane + off] = (uint16_t)(t3 >> 16);
    }
  }
  for (int t = tid; t < n * k; t += 128) {
    const int r = t % n, kk = t / n;
    const int off = b_mn ? tc::plane_offset(k, kk, r) : tc::plane_offset(n, r, kk);
    if constexpr (F16) {
      uint32_t hi, lo;
      tc::split_f16x2_pair(B[r * k + kk], 0.f, hi, lo);
      Bp[off] = (uint16_t)hi, Bp[b_plane + off] = (uint16_t)lo;
    } else {
      uint32_t t1, t2, t3;
      tc::split_bf16x3(B[r * k + kk], t1, t2, t3);
      Bp[off] = (uint16_t)(t1 >> 16), Bp[b_plane + off] = (uint16_t)(t2 >> 16), Bp[2 * b_plane + off] = (uint16_t)(t3 >> 16);
    }
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
  const uint32_t tmem_a = tmem_slot + (uint32_t)n;  // A operand: 3 planes of k / 2 columns
  const uint32_t a_cols = (uint32_t)k / 2;

  if (a_tmem) {  // thread = row of A
    for (int k0 = 0; k0 < k; k0 += 32) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = A[tid * k + k0 + i];
      if constexpr (F16)
        tc::tmem_st_row32_f16x2(tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)k0 / 2, a_cols, v);
      else
        tc::tmem_st_row32_bf16x3(tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)k0 / 2, a_cols, v);
    }
  }
  if (init_d) {
    for (int c0 = 0; c0 < n; c0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = D[tid * n + c0 + i];
      tc::tmem_st16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    }
  }
  if (init_d || a_tmem) {
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }

  long long t0 = 0;
  if (warp == 0) {
    if (tc::elect_one_sync()) {
      const uint32_t idesc = F16 ? tc::idesc_f16(128, n, a_mn, b_mn) : tc::idesc_bf16(128, n, a_mn, b_mn);
      const uint64_t ad0 = a_mn ? tc::plane_desc_mn_base(smem_u32(Ap), a_rows) : tc::plane_desc_k_base(smem_u32(Ap), a_rows);
      const uint64_t bd0 = b_mn ? tc::plane_desc_mn_base(smem_u32(Bp), b_rows) : tc::plane_desc_k_base(smem_u32(Bp), b_rows);
      const uint32_t a_step = a_mn ? tc::kPlaneMnStepBytes : tc::plane_k_step_bytes(a_rows);
      const uint32_t b_step = b_mn ? tc::kPlaneMnStepBytes : tc::plane_k_step_bytes(b_rows);
      bool acc = init_d;
      t0 = clock64();
      auto one = [&](int pass, int ks) {
        const int ta = F16 ? tc::f16x2_term_a(pass) : tc::bf16x3_term_a(pass), tb = F16 ? tc::f16x2_term_b(pass) : tc::bf16x3_term_b(pass);
        const uint64_t ad = tc::desc_advance(ad0, (uint32_t)(ta * a_plane * 2) + ks * a_step);
        const uint64_t bd = tc::desc_advance(bd0, (uint32_t)(tb * b_plane * 2) + ks * b_step);
        if (a_tmem)
          tc::mma_bf16_ts(tmem_d, tmem_a + (uint32_t)ta * a_cols + (uint32_t)ks * 8, bd, idesc, acc);
        else
          tc::mma_bf16(tmem_d, ad, bd, idesc, acc);
        acc = true;
      };
      if constexpr (KC > 0) {
#pragma unroll
        for (int pass = 0; pass < (F16 ? 3 : 6); ++pass)
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) one(pass, ks);
      } else {
        for (int pass = 0; pass < (F16 ? 3 : 6); ++pass)
          for (int ks = 0; ks < k / 16; ++ks) one(pass, ks);
      }
      tc::mma_commit(&bar);
      if (dbg) dbg[0] = clock64() - t0;  // issue time
    }
    __syncwarp();
  }
  const bool ok = tc::mbar_wait_bounded(&bar, 0);
  tc::fence_after_sync();
  if (t0 != 0 && dbg) dbg[1] = clock64() - t0;  // issue -> completion of 6 * k / 16 MMAs
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
  B200_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 8 && k % 8 == 0 && (passes == 1 || passes == 2 || passes == 3 || passes == 6) &&
                   flags >= 0 && flags < 16,
               B200_E_INVALID,
               "b200_tc_gemm_selftest: need 16 <= n <= 256 (multiple of 16), k multiple of 8, passes in {1, 2, 3, 6}, flags < 16");
  uint32_t cols = 32;
  while ((int)cols < n) cols <<= 1;
  if (passes == 6 || passes == 2) {
    B200_REQUIRE(k % 16 == 0, B200_E_INVALID, "b200_tc_gemm_selftest: bf16 operands need k %% 16 == 0");
    if (flags & 8) {
      B200_REQUIRE(k % 32 == 0 && !(flags & 1) && n + 3 * k / 2 <= 512, B200_E_INVALID,
                   "b200_tc_gemm_selftest: A in tensor memory needs k %% 32 == 0, flag bit 0 clear, n + 1.5 k <= 512 columns");
      while ((int)cols < n + 3 * k / 2) cols <<= 1;
    }
    const size_t a_plane = (flags & 1) ? tc::plane_halves(k, 128) : tc::plane_halves(128, k);
    const size_t b_plane = (flags & 2) ? tc::plane_halves(k, n) : tc::plane_halves(n, k);
    // the M = 128 / N = n reads of a K-major plane with a partial last row group stay inside the plane; MN-major
    // planes are read exactly
    const size_t smem = sizeof(uint16_t) * 3 * (a_plane + b_plane);
    B200_REQUIRE(smem <= 200 * 1024, B200_E_UNSUPPORTED, "b200_tc_gemm_selftest: operands need %zu bytes of shared memory", smem);
    auto kern = (passes == 2) ? ((k == 64) ? tc_gemm_selftest_bf16_kernel<64, true> : tc_gemm_selftest_bf16_kernel<0, true>)
                              : ((k == 64) ? tc_gemm_selftest_bf16_kernel<64, false> : tc_gemm_selftest_bf16_kernel<0, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "tc selftest smem attribute");
    kern<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d, n, k, flags, cols, status, tc_debug_buffer());
    B200_CHECK_LAUNCH("tc_gemm_selftest_bf16_kernel");
    return B200_OK;
  }
  B200_REQUIRE((flags & 11) == 0, B200_E_UNSUPPORTED,
               "b200_tc_gemm_selftest: MN-major / tensor-memory operands are implemented for passes = 6 only (tf32 operands "
               "cannot be read MN-major from the no-swizzle layout)");
  const size_t smem = sizeof(float) * 2 * (tc::operand_floats(128, k) + tc::operand_floats(n, k));
  B200_REQUIRE(smem <= 200 * 1024, B200_E_UNSUPPORTED, "b200_tc_gemm_selftest: operands need %zu bytes of shared memory", smem);
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "tc selftest smem attribute");
  tc_gemm_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d, n, k, passes, flags, cols, status);
  B200_CHECK_LAUNCH("tc_gemm_selftest_kernel");
  return B200_OK;
}
