// Fused LocSE + attentive pooling FORWARD with the attention contraction on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, 3xTF32 split, accumulator in TMEM) -- the c in {64, 128} levels of
// LocalFeatureAggregation (myria3d/models/modules/pyg_randla_net.py:121-152).
//
// Orientation: D[n][e] = sum_m W_att[n][m] * F[e][m]  -- output CHANNELS on the 128 TMEM lanes, the tile's EDGES
// on the TMEM columns.  A thread of the epilogue therefore owns one channel and reads, per centre, 16 consecutive
// columns = the 16 neighbours of that centre: the neighbourhood softmax and the weighted sum stay thread-local
// register reductions exactly as in the FMA kernel (lfa.cu), with no shuffles and no atomics.
//
//   operands (tc.cuh layout, tf32 hi/lo pairs):  A = W_att [128 (rows >= c zero) x c], resident for the whole CTA;
//                                                B = F tile [NE edges x c], rebuilt per tile by all 256 threads
//   per tile: build F (gather x_j as float4 = one 16-byte k-chunk, encoder lrelu(enc_w.q + enc_b) 4 channels at
//             a time) -> fence -> one thread issues 3 * c/8 MMAs + commit -> mbarrier wait -> epilogue
//             (tcgen05.ld 16 columns per centre, exp, weighted sum with f = hi + lo from shared memory).
#include <math_constants.h>
#include <stdlib.h>

#include "tc.cuh"

namespace b200 {

constexpr int LTC_THREADS = 256;
constexpr int LTC_KT = 16;

template <int C, int NE>
__global__ void __launch_bounds__(LTC_THREADS, 1)
lfa_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pos, const int32_t* __restrict__ nbr,
                  const float* __restrict__ enc_w, const float* __restrict__ enc_b,
                  const float* __restrict__ att_wt /* [m][n] = W_att[n][m], as the C ABI passes it */,
                  float* __restrict__ out, int64_t n, int64_t ntiles, uint32_t tmem_cols) {
  constexpr int H = C / 2, TC = NE / LTC_KT, H4 = H / 4;
  static_assert(C % 8 == 0 && C <= 128 && NE % 16 == 0 && NE <= 256 && TC % 2 == 0, "tile shape");
  extern __shared__ __align__(128) float ltc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr size_t W_FLOATS = tc::operand_floats(128, C), F_FLOATS = tc::operand_floats(NE, C);
  float* Wh = ltc_smem;
  float* Wl = Wh + W_FLOATS;
  float* Fh = Wl + W_FLOATS;
  float* Fl = Fh + F_FLOATS;
  float4* Q = reinterpret_cast<float4*>(Fl + F_FLOATS);  // (p_j, dist) per edge
  float4* P = Q + NE;                                      // p_i per centre
  int* NB = reinterpret_cast<int*>(P + TC);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- resident A operand: element (row n, k = m) = W_att[n][m] = att_wt[m][n], tf32 hi/lo, zero rows n >= c
  for (int idx = tid; idx < (128 - C) * (C / 4); idx += LTC_THREADS) {  // padding rows (c = 64 only)
    const int row = C + idx / (C / 4), m4 = idx % (C / 4);
    const int off = (m4 * 129 + row) * 4;
    *reinterpret_cast<float4*>(Wh + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(Wl + off) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int idx = tid; idx < C * (C / 4); idx += LTC_THREADS) {
    const int mm = idx / (C / 4), n4 = idx % (C / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(att_wt + (int64_t)mm * C) + n4);
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float hi, lo;
      tc::split_tf32(e[u], hi, lo);
      const int off = tc::operand_offset(128, 4 * n4 + u, mm);
      Wh[off] = hi;
      Wl[off] = lo;
    }
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = tc::idesc_tf32(128, NE);
  uint32_t phase = 0;
  bool alive = true;

  for (int64_t tile = blockIdx.x; tile < ntiles && alive; tile += gridDim.x) {
    const int64_t tile_base = tile * TC;
    // ---- neighbour ids and geometry
    for (int e = tid; e < NE; e += LTC_THREADS) {
      const int g = e / LTC_KT, kk = e % LTC_KT;
      const int64_t i = tile_base + g;
      int j = -1;
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), pv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n) {
        j = __ldg(nbr + i * LTC_KT + kk);
        pv = make_float4(__ldg(pos + 3 * i), __ldg(pos + 3 * i + 1), __ldg(pos + 3 * i + 2), 0.f);
        if (j >= 0) {
          const float pjx = __ldg(pos + 3 * (int64_t)j), pjy = __ldg(pos + 3 * (int64_t)j + 1), pjz = __ldg(pos + 3 * (int64_t)j + 2);
          const float dx = pjx - pv.x, dy = pjy - pv.y, dz = pjz - pv.z;
          qv = make_float4(pjx, pjy, pjz, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
        }
      }
      NB[e] = j;
      Q[e] = qv;
      if (kk == 0) P[g] = pv;
    }
    __syncthreads();
    // ---- B operand, first half of K: gathered neighbour features (one float4 = one 16-byte k-chunk)
    for (int t = tid; t < NE * H4; t += LTC_THREADS) {
      const int e = t / H4, m4 = t % H4;
      const int j = NB[e];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j >= 0) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)j * H) + m4);
      float4 hi, lo;
      tc::split_tf32(v.x, hi.x, lo.x), tc::split_tf32(v.y, hi.y, lo.y);
      tc::split_tf32(v.z, hi.z, lo.z), tc::split_tf32(v.w, hi.w, lo.w);
      const int off = (m4 * (NE + 1) + e) * 4;
      *reinterpret_cast<float4*>(Fh + off) = hi;
      *reinterpret_cast<float4*>(Fl + off) = lo;
    }
    // ---- second half of K: local spatial encoding, 4 channels per thread and edge
    for (int t = tid; t < NE * H4; t += LTC_THREADS) {
      const int e = t / H4, c4 = t % H4;
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (NB[e] >= 0) {
        const float4 p = P[e / LTC_KT], q = Q[e];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* w = enc_w + (4 * c4 + u) * 7;
          float s = __ldg(enc_b + 4 * c4 + u);
          s = fmaf(__ldg(w + 0), p.x, s), s = fmaf(__ldg(w + 1), p.y, s), s = fmaf(__ldg(w + 2), p.z, s);
          s = fmaf(__ldg(w + 3), q.x, s), s = fmaf(__ldg(w + 4), q.y, s), s = fmaf(__ldg(w + 5), q.z, s);
          s = fmaf(__ldg(w + 6), q.w, s);
          z[u] = lrelu(s, kLReluSlope);
        }
      }
      float4 hi, lo;
      tc::split_tf32(z[0], hi.x, lo.x), tc::split_tf32(z[1], hi.y, lo.y);
      tc::split_tf32(z[2], hi.z, lo.z), tc::split_tf32(z[3], hi.w, lo.w);
      const int off = ((H4 + c4) * (NE + 1) + e) * 4;
      *reinterpret_cast<float4*>(Fh + off) = hi;
      *reinterpret_cast<float4*>(Fl + off) = lo;
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    // ---- a[n][e] on the tensor cores: hi*hi + lo*hi + hi*lo
    if (tid == 0) {
      const uint32_t lbo_a = tc::lbo_bytes(128), lbo_b = tc::lbo_bytes(NE);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_base = smem_u32((pass == 1) ? Wl : Wh);
        const uint32_t b_base = smem_u32((pass == 2) ? Fl : Fh);
#pragma unroll 4
        for (int ks = 0; ks < C / 8; ++ks) {
          const uint64_t ad = tc::smem_desc(a_base + (uint32_t)(2 * ks) * lbo_a, lbo_a, tc::kSboBytes);
          const uint64_t bd = tc::smem_desc(b_base + (uint32_t)(2 * ks) * lbo_b, lbo_b, tc::kSboBytes);
          tc::mma_tf32(tmem_d, ad, bd, idesc, (pass | ks) != 0);
        }
      }
      tc::mma_commit(&bar);
    }
    alive = tc::mbar_wait_bounded(&bar, phase);
    phase ^= 1u;
    alive = __syncthreads_and(alive) != 0;
    tc::fence_after_sync();
    if (!alive) break;

    // ---- epilogue: thread = channel (TMEM lane); warps w and w+4 share a lane quarter and split the centres
    {
      const int lq = warp & 3, half = warp >> 2;
      const int ch = lq * 32 + lane;
#pragma unroll 1
      for (int g = half * (TC / 2); g < (half + 1) * (TC / 2); ++g) {
        float a[16];
        tc::tmem_ld16(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(g * LTC_KT), a);
        const int64_t i = tile_base + g;
        if (ch < C && i < n) {
          int deg = 0;
#pragma unroll
          for (int k = 0; k < LTC_KT; ++k) deg += (NB[g * LTC_KT + k] >= 0) ? 1 : 0;
          float mx = -CUDART_INF_F;
#pragma unroll
          for (int k = 0; k < LTC_KT; ++k)
            if (k < deg) mx = fmaxf(mx, a[k]);
          float sum = 0.f, o = 0.f;
          const int foff = ((ch >> 2) * (NE + 1) + g * LTC_KT) * 4 + (ch & 3);
#pragma unroll
          for (int k = 0; k < LTC_KT; ++k)
            if (k < deg) {
              const float p = __expf(a[k] - mx);
              const float f = Fh[foff + 4 * k] + Fl[foff + 4 * k];  // hi + lo == the fp32 feature, exactly
              sum += p;
              o = fmaf(p, f, o);
            }
          out[i * C + ch] = o / (sum + 1e-16f);
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();  // TMEM and the F tile are free for the next tile
    tc::fence_after_sync();
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

template <int C, int NE>
static int launch_lfa_tc_fwd(const float* x, const float* pos, const int32_t* nbr, const float* enc_w, const float* enc_b,
                             const float* att_wt, float* out, int64_t n, cudaStream_t st) {
  constexpr int TC = NE / LTC_KT;
  const size_t smem = sizeof(float) * (2 * tc::operand_floats(128, C) + 2 * tc::operand_floats(NE, C)) +
                      sizeof(float4) * (NE + TC) + sizeof(int) * NE;
  auto kern = lfa_tc_fwd_kernel<C, NE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "lfa_tc_fwd smem attribute");
  const int64_t ntiles = ceil_div(n, TC);
  int64_t grid = num_sms();
  if (grid > ntiles) grid = ntiles;
  uint32_t cols = 32;
  while ((int)cols < NE) cols <<= 1;
  kern<<<(unsigned)grid, LTC_THREADS, smem, st>>>(x, pos, nbr, enc_w, enc_b, att_wt, out, n, ntiles, cols);
  B200_CHECK_LAUNCH("lfa_tc_fwd_kernel");
  return B200_OK;
}

// returns B200_E_UNSUPPORTED when this (c, kt) has no tensor-core kernel (the caller then uses the FMA kernel)
int lfa_tc_fwd_dispatch(const float* x, const float* pos, const int32_t* nbr, const float* enc_w, const float* enc_b,
                        const float* att_wt, float* out, int64_t n, int c, int kt, cudaStream_t st) {
  // Opt-in (B200_LFA_TCGEN05=1): numerically validated against the oracle, but with one CTA per SM and no overlap
  // between the gather/build phase and the MMA it is still slower than the 3-CTA/SM FMA kernel.
  const char* opt = getenv("B200_LFA_TCGEN05");
  if (kt != LTC_KT || !tensor_cores_enabled() || !(opt && opt[0] == '1')) return B200_E_UNSUPPORTED;
  if (c == 64) return launch_lfa_tc_fwd<64, 128>(x, pos, nbr, enc_w, enc_b, att_wt, out, n, st);
  if (c == 128) return launch_lfa_tc_fwd<128, 64>(x, pos, nbr, enc_w, enc_b, att_wt, out, n, st);
  return B200_E_UNSUPPORTED;
}

}  // namespace b200
