// Fused LocSE + attentive pooling, FORWARD and BACKWARD, with every contraction on the 5th-generation tensor cores
// (tcgen05.mma kind::f16 on bf16 x 3 split operands = fp32-grade products, accumulators in TMEM) -- the
// c in {32, 64, 128} levels of LocalFeatureAggregation (myria3d/models/modules/pyg_randla_net.py:121-152).
// Production path of b200_lfa_fwd / b200_lfa_bwd for these widths; narrower levels (c = 8, 16: K = 8 / 16
// contractions) and c = 256 stay on the FMA kernels of lfa.cu.
//
// Orientation: CHANNELS on the 128 TMEM lanes, the tile's EDGES on the TMEM columns.  An epilogue thread owns one
// channel and reads, per centre, KT consecutive columns = the neighbours of that centre: the neighbourhood softmax
// and the weighted sum are thread-local register reductions (no shuffles, no atomics), exactly like the FMA kernel.
//
// One tile = NE edges = NE / KT centres.  Shared-memory operands are planes of bf16 (tc.cuh: 16-byte vectors of 8
// elements along the "chunked" dimension, rows 16 bytes apart), three planes (t1, t2, t3 with v = t1 + t2 + t3) each:
//     W  rows = output channel n, chunked along m     resident for the whole CTA     (W_att[n][m])
//     F  rows = edge e,           chunked along m     rebuilt per tile  (x_j gather | lrelu(enc_w . q + enc_b))
//     dA rows = channel n,        chunked along e     backward only: softmax gradient, written by the epilogue threads
// and every plane is read BOTH ways by the tensor core -- K-major (rows = M/N index) and MN-major (rows = K index) --
// so that no operand is ever transposed or copied (16-bit operands only: see tc.cuh):
//     MMA1  S [n][e]  = sum_m W[n][m]  F[e][m]      A = W  (K-major),  B = F  (K-major)      scores
//     MMA3  dF[m][e] += sum_n W[n][m] dA[e][n]      A = W  (MN-major), B = dA (MN-major)     accumulator pre-loaded
//                                                                                             with s*go by tcgen05.st
//     MMA4  dW[n][m] += sum_e dA[e][n] F[e][m]      A = dA (K-major),  B = F  (MN-major)     TMEM-resident across ALL
//                                                                                             tiles of the CTA
// The attention-weight gradient therefore never leaves the SM until the CTA is done (one red.global per element and
// CTA); round 1 streamed two [E, c] tensors through HBM for it (370 MB at c = 64).
//
// Per tile (backward): build F -> MMA1 -> E1: softmax, o, dA -> smem, s*go -> TMEM -> MMA3 + MMA4 -> E3: encoder
// gradients in registers (carried across tiles), x-gradients transposed through shared memory and scattered with
// 16-byte vector reductions.  The epilogue threads take the fp32 feature values they need (weighted sum, f - o) from
// where they are exact: x_j straight from global memory (32 lanes = 32 consecutive channels = one 128-byte line,
// L1-resident from the build) and the encoding recomputed from the 7 geometry numbers.
#include <math_constants.h>

#include "tc.cuh"

namespace b200 {

constexpr int LTC_THREADS = 256;
long long* tc_debug_buffer();  // runtime.cu

template <int C, int NE, int KT, bool BWD>
struct LtcPlan {
  static constexpr int H = C / 2, H8 = H / 8, TC = NE / KT;
  static constexpr size_t W_PLANE = tc::plane_halves(C, C) * 2;     // bytes: [C/8][C+1] x 16
  static constexpr size_t F_PLANE = tc::plane_halves(NE, C) * 2;    //        [C/8][NE+1] x 16
  static constexpr size_t DA_PLANE = tc::plane_halves(C, NE) * 2;   //        [NE/8][C+1] x 16
  static constexpr size_t T_BYTES = tc::operand_floats(NE, H) * 4;  // fp32 [H/4][NE+1][4]: x-gradient transposition
  // buffer order: W | dA | F | T | Q P EW | NB DEG.  The M = 128 reads of planes with fewer than 128 rows / 16
  // chunks run past their end into the NEXT buffers (garbage in TMEM lanes >= C that nobody reads); F (read
  // exactly) comes last among the operands so that every such over-read stays inside the allocation.
  static constexpr size_t OFF_W = 0;
  static constexpr size_t OFF_DA = OFF_W + 3 * W_PLANE;
  static constexpr size_t OFF_F = OFF_DA + (BWD ? 3 * DA_PLANE : 0);
  static constexpr size_t OFF_T = OFF_F + 3 * F_PLANE;
  static constexpr size_t OFF_Q = OFF_T + (BWD ? T_BYTES : 0);
  static constexpr size_t OFF_P = OFF_Q + 16 * (size_t)NE;
  static constexpr size_t OFF_EW = OFF_P + 16 * (size_t)TC;
  static constexpr size_t OFF_NB = OFF_EW + 32 * (size_t)H;
  static constexpr size_t OFF_DEG = OFF_NB + 4 * (size_t)NE;
  static constexpr size_t SMEM_BYTES = OFF_DEG + 16 * ((TC + 3) / 4);
  static constexpr uint32_t TMEM_COLS_NEEDED = NE + (BWD ? C : 0);
  static constexpr uint32_t TMEM_COLS =
      TMEM_COLS_NEEDED <= 32 ? 32 : TMEM_COLS_NEEDED <= 64 ? 64 : TMEM_COLS_NEEDED <= 128 ? 128 : TMEM_COLS_NEEDED <= 256 ? 256 : 512;
  static_assert(C % 32 == 0 && C <= 128, "channels map to whole warps of TMEM lanes");
  static_assert(KT == 16 || KT == 32, "neighbour slots");
  static_assert(NE % 32 == 0 && NE <= 256 && TC >= 2 && TC % 2 == 0, "tile shape");
  // worst over-reads, measured from the start of the LAST plane of each buffer
  static_assert(16 * (size_t)(C + 1) * 16 + 2048 <= SMEM_BYTES - (OFF_W + 2 * W_PLANE), "W over-read");
  static_assert(!BWD || (size_t)(NE / 8 * (C + 1) + 128) * 16 <= SMEM_BYTES - (OFF_DA + 2 * DA_PLANE), "dA over-read");
};

__device__ __forceinline__ void ltc_wait(uint64_t* bar, uint32_t parity) {
  if (!tc::mbar_wait_bounded(bar, parity, 20000000u)) __trap();  // a wrong descriptor must not hang the GPU
}

// 8 fp32 values -> three 16-byte vectors of bf16 terms
__device__ __forceinline__ void split8_bf16x3(const float (&v)[8], uint4& p1, uint4& p2, uint4& p3) {
  uint32_t t1[8], t2[8], t3[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) tc::split_bf16x3(v[u], t1[u], t2[u], t3[u]);
  p1 = make_uint4(tc::pack_hi16(t1[0], t1[1]), tc::pack_hi16(t1[2], t1[3]), tc::pack_hi16(t1[4], t1[5]), tc::pack_hi16(t1[6], t1[7]));
  p2 = make_uint4(tc::pack_hi16(t2[0], t2[1]), tc::pack_hi16(t2[2], t2[3]), tc::pack_hi16(t2[4], t2[5]), tc::pack_hi16(t2[6], t2[7]));
  p3 = make_uint4(tc::pack_hi16(t3[0], t3[1]), tc::pack_hi16(t3[2], t3[3]), tc::pack_hi16(t3[4], t3[5]), tc::pack_hi16(t3[6], t3[7]));
}

template <int C, int NE, int KT, bool BWD, int MINB>
__global__ void __launch_bounds__(LTC_THREADS, MINB)
lfa_tc_kernel(const float* __restrict__ x, const float* __restrict__ pos, const int32_t* __restrict__ nbr,
              const float* __restrict__ enc_w, const float* __restrict__ enc_b,
              const float* __restrict__ att_w /* backward: W_att [n][m]; forward: att_wt [m][n] */,
              float* __restrict__ out,                                                   // forward
              const float* __restrict__ grad_out, float* __restrict__ grad_x,            // backward
              float* __restrict__ grad_enc_w, float* __restrict__ grad_enc_b, float* __restrict__ grad_att_w,
              int64_t n, int64_t ntiles, long long* __restrict__ dbg) {
  using Plan = LtcPlan<C, NE, KT, BWD>;
  constexpr int H = Plan::H, H8 = Plan::H8, TC = Plan::TC;
  extern __shared__ __align__(128) unsigned char ltc_smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  unsigned char* Wp = ltc_smem + Plan::OFF_W;    // 3 planes of W_PLANE bytes
  unsigned char* dAp = ltc_smem + Plan::OFF_DA;  // 3 planes of DA_PLANE bytes
  unsigned char* Fp = ltc_smem + Plan::OFF_F;    // 3 planes of F_PLANE bytes
  float* T = reinterpret_cast<float*>(ltc_smem + Plan::OFF_T);
  float4* Q = reinterpret_cast<float4*>(ltc_smem + Plan::OFF_Q);    // (p_j, dist) per edge
  float4* P = reinterpret_cast<float4*>(ltc_smem + Plan::OFF_P);    // p_i per centre
  float4* EW = reinterpret_cast<float4*>(ltc_smem + Plan::OFF_EW);  // [H][2]: (w0..w3), (w4, w5, w6, bias)
  int* NB = reinterpret_cast<int*>(ltc_smem + Plan::OFF_NB);
  int* DEG = reinterpret_cast<int*>(ltc_smem + Plan::OFF_DEG);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int nts = 0;
#define LTC_TS()                                                                            \
  do {                                                                                      \
    if (dbg != nullptr && blockIdx.x == 0 && tid == 0 && nts < 120) dbg[nts++] = clock64(); \
  } while (0)
  LTC_TS();

  // ---- resident W operand: element (row n, position m) = W_att[n][m] as bf16 x 3.  The backward gets att_w = W_att
  // row-major; the forward ABI only carries att_wt[m][n] = W_att[n][m] (one strided pass per CTA).
  for (int idx = tid; idx < C * (C / 8); idx += LTC_THREADS) {
    const int nn = idx % C, j = idx / C;  // consecutive lanes = consecutive rows: conflict-free 16-byte stores
    float v[8];
    if constexpr (BWD) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(att_w + (int64_t)nn * C + 8 * j));
      const float4 b = __ldg(reinterpret_cast<const float4*>(att_w + (int64_t)nn * C + 8 * j + 4));
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(att_w + (int64_t)(8 * j + u) * C + nn);
    }
    uint4 p1, p2, p3;
    split8_bf16x3(v, p1, p2, p3);
    const size_t off = ((size_t)j * (C + 1) + nn) * 16;
    *reinterpret_cast<uint4*>(Wp + off) = p1;
    *reinterpret_cast<uint4*>(Wp + Plan::W_PLANE + off) = p2;
    *reinterpret_cast<uint4*>(Wp + 2 * Plan::W_PLANE + off) = p3;
  }
  for (int idx = tid; idx < H; idx += LTC_THREADS) {
    const float* w = enc_w + idx * 7;
    EW[2 * idx] = make_float4(__ldg(w), __ldg(w + 1), __ldg(w + 2), __ldg(w + 3));
    EW[2 * idx + 1] = make_float4(__ldg(w + 4), __ldg(w + 5), __ldg(w + 6), __ldg(enc_b + idx));
  }

  if (warp == 0) tc::tmem_alloc(&tmem_slot, Plan::TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_s = tmem_slot;                  // scores S / dF accumulator: NE columns
  const uint32_t tmem_dw = tmem_slot + (uint32_t)NE;  // dW accumulator: C columns (backward)
  uint32_t phase = 0;

  // epilogue role: TMEM lane quadrant (= 32 channels) and the half of the tile's centres this warp handles
  const int lq = warp & 3, half = warp >> 2;
  const int ch = lq * 32 + lane;
  const bool ch_active = (lq * 32) < C;  // warp-uniform
  const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
  // encoder weights of this thread's channel (used when ch >= H)
  float4 mw0 = make_float4(0.f, 0.f, 0.f, 0.f), mw1 = mw0;
  if (ch_active && ch >= H) {
    mw0 = EW[2 * (ch - H)];
    mw1 = EW[2 * (ch - H) + 1];
  }

  float gw[8];  // encoder-gradient partials of channel ch (>= H), carried across tiles
#pragma unroll
  for (int t = 0; t < 8; ++t) gw[t] = 0.f;
  bool first_tile = true;
  LTC_TS();

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t tile_base = tile * TC;
    // ---- B1: neighbour ids, geometry, degrees
    for (int e = tid; e < NE; e += LTC_THREADS) {
      const int g = e / KT, kk = e % KT;
      const int64_t i = tile_base + g;
      int j = -1;
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), pv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n) {
        j = __ldg(nbr + i * KT + kk);
        pv = make_float4(__ldg(pos + 3 * i), __ldg(pos + 3 * i + 1), __ldg(pos + 3 * i + 2), 0.f);
        if (j >= 0) {
          const float pjx = __ldg(pos + 3 * (int64_t)j), pjy = __ldg(pos + 3 * (int64_t)j + 1),
                      pjz = __ldg(pos + 3 * (int64_t)j + 2);
          const float dx = pjx - pv.x, dy = pjy - pv.y, dz = pjz - pv.z;
          qv = make_float4(pjx, pjy, pjz,
                           sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
        }
      }
      NB[e] = j;
      Q[e] = qv;
      const unsigned valid = __ballot_sync(0xffffffffu, j >= 0);  // NE % 32 == 0: whole warps in this loop
      if (kk == 0) {
        P[g] = pv;
        DEG[g] = (KT == 32) ? __popc(valid) : __popc((valid >> (lane & 16)) & 0xffffu);
      }
    }
    __syncthreads();
    LTC_TS();
    // ---- B2: F[:, 0:H) = gathered neighbour features, 8 channels (one 16-byte bf16 vector per term) per item
#pragma unroll
    for (int t0 = 0; t0 < NE * H8; t0 += LTC_THREADS) {
      const int t = t0 + tid;
      if ((NE * H8) % LTC_THREADS == 0 || t < NE * H8) {
        const int e = t / H8, m8 = t % H8;
        const int j = NB[e];
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (j >= 0) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(x + (int64_t)j * H + 8 * m8));
          const float4 b = __ldg(reinterpret_cast<const float4*>(x + (int64_t)j * H + 8 * m8 + 4));
          v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
        }
        uint4 p1, p2, p3;
        split8_bf16x3(v, p1, p2, p3);
        const size_t off = ((size_t)m8 * (NE + 1) + e) * 16;
        *reinterpret_cast<uint4*>(Fp + off) = p1;
        *reinterpret_cast<uint4*>(Fp + Plan::F_PLANE + off) = p2;
        *reinterpret_cast<uint4*>(Fp + 2 * Plan::F_PLANE + off) = p3;
      }
    }
    // ---- B3: F[:, H:C) = local spatial encoding, 8 channels per item
#pragma unroll
    for (int t0 = 0; t0 < NE * H8; t0 += LTC_THREADS) {
      const int t = t0 + tid;
      if ((NE * H8) % LTC_THREADS == 0 || t < NE * H8) {
        const int e = t / H8, c8 = t % H8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (NB[e] >= 0) {
          const float4 p = P[e / KT], q = Q[e];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 w0 = EW[2 * (8 * c8 + u)], w1 = EW[2 * (8 * c8 + u) + 1];
            float s = fmaf(w0.z, p.z, fmaf(w0.y, p.y, fmaf(w0.x, p.x, w1.w)));
            s = fmaf(w0.w, q.x, s), s = fmaf(w1.x, q.y, s), s = fmaf(w1.y, q.z, s);
            s = fmaf(w1.z, q.w, s);
            v[u] = lrelu(s, kLReluSlope);
          }
        }
        uint4 p1, p2, p3;
        split8_bf16x3(v, p1, p2, p3);
        const size_t off = ((size_t)(H8 + c8) * (NE + 1) + e) * 16;
        *reinterpret_cast<uint4*>(Fp + off) = p1;
        *reinterpret_cast<uint4*>(Fp + Plan::F_PLANE + off) = p2;
        *reinterpret_cast<uint4*>(Fp + 2 * Plan::F_PLANE + off) = p3;
      }
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    LTC_TS();

    // ---- MMA1: S[n][e] = sum_m W[n][m] F[e][m]   (six bf16 cross products)
    if (tid == 0) {
      constexpr uint32_t idesc = tc::idesc_bf16(128, NE);
#pragma unroll
      for (int pass = 0; pass < 6; ++pass) {
        const uint32_t a_base = smem_u32(Wp + tc::bf16x3_term_a(pass) * Plan::W_PLANE);
        const uint32_t b_base = smem_u32(Fp + tc::bf16x3_term_b(pass) * Plan::F_PLANE);
#pragma unroll
        for (int k0 = 0; k0 < C; k0 += 16)
          tc::mma_bf16(tmem_s, tc::plane_desc_k(a_base, C, k0), tc::plane_desc_k(b_base, NE, k0), idesc, (pass | k0) != 0);
      }
      tc::mma_commit(&bars[0]);
    }
    ltc_wait(&bars[0], phase);
    tc::fence_after_sync();
    LTC_TS();

    // ---- E1: thread = channel; softmax over the neighbourhood, pooled output, (backward) softmax gradient
    if (ch_active) {
#pragma unroll 1
      for (int g = half * (TC / 2); g < (half + 1) * (TC / 2); ++g) {
        float a[KT], f[KT];
#pragma unroll
        for (int k0 = 0; k0 < KT; k0 += 16) {
          float v[16];
          tc::tmem_ld16(tmem_s + lane_base + (uint32_t)(g * KT + k0), v);
#pragma unroll
          for (int k = 0; k < 16; ++k) a[k0 + k] = v[k];
        }
        const int deg = DEG[g];
        const int64_t i = tile_base + g;
        // the fp32 feature of (edge, this channel): gathered x_j or the recomputed encoding (same operation order as
        // the build: bit-identical to the value the tensor core saw, before the bf16 x 3 split)
        if (ch < H) {
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const int j = NB[g * KT + k];
            f[k] = (j >= 0) ? __ldg(x + (int64_t)j * H + ch) : 0.f;
          }
        } else {
          const float4 p = P[g];
          const float sp = fmaf(mw0.z, p.z, fmaf(mw0.y, p.y, fmaf(mw0.x, p.x, mw1.w)));
#pragma unroll
          for (int k = 0; k < KT; ++k) {
            const float4 q = Q[g * KT + k];
            float s = sp;
            s = fmaf(mw0.w, q.x, s), s = fmaf(mw1.x, q.y, s), s = fmaf(mw1.y, q.z, s);
            s = fmaf(mw1.z, q.w, s);
            f[k] = (k < deg) ? lrelu(s, kLReluSlope) : 0.f;
          }
        }
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < deg) mx = fmaxf(mx, a[k]);
        float sum = 0.f, o = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const float p = (k < deg) ? __expf(a[k] - mx) : 0.f;
          a[k] = p;
          sum += p;
          o = fmaf(p, f[k], o);
        }
        const float inv = 1.f / (sum + 1e-16f);
        o *= inv;
        if constexpr (!BWD) {
          if (i < n) out[i * C + ch] = o;
        } else {
          const float gi = (i < n) ? inv * __ldg(grad_out + i * C + ch) : 0.f;
#pragma unroll
          for (int k8 = 0; k8 < KT; k8 += 8) {
            float da[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float sg = a[k8 + u] * gi;  // s * go: direct term of dF (0 for invalid edges: p = 0)
              da[u] = sg * (f[k8 + u] - o);     // gradient w.r.t. the score
              a[k8 + u] = sg;
            }
            uint4 p1, p2, p3;
            split8_bf16x3(da, p1, p2, p3);
            const size_t off = ((size_t)((g * KT + k8) >> 3) * (C + 1) + ch) * 16;
            *reinterpret_cast<uint4*>(dAp + off) = p1;
            *reinterpret_cast<uint4*>(dAp + Plan::DA_PLANE + off) = p2;
            *reinterpret_cast<uint4*>(dAp + 2 * Plan::DA_PLANE + off) = p3;
          }
#pragma unroll
          for (int k0 = 0; k0 < KT; k0 += 16) {
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = a[k0 + k];
            tc::tmem_st16(tmem_s + lane_base + (uint32_t)(g * KT + k0), v);  // dF accumulator <- s * go
          }
        }
      }
      if constexpr (BWD) tc::tmem_st_wait();
    }
    if constexpr (!BWD) {
      tc::fence_before_sync();
      __syncthreads();  // TMEM and the F tile are free for the next tile
      tc::fence_after_sync();
      LTC_TS();
    } else {
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      LTC_TS();

      // ---- MMA3: dF[m][e] += sum_n W[n][m] dA[e][n];  MMA4: dW[n][m] += sum_e dA[e][n] F[e][m]
      if (tid == 0) {
        {
          constexpr uint32_t idesc = tc::idesc_bf16(128, NE, /*a_mn=*/true, /*b_mn=*/true);
#pragma unroll
          for (int pass = 0; pass < 6; ++pass) {
            const uint32_t a_base = smem_u32(Wp + tc::bf16x3_term_a(pass) * Plan::W_PLANE);
            const uint32_t b_base = smem_u32(dAp + tc::bf16x3_term_b(pass) * Plan::DA_PLANE);
#pragma unroll
            for (int k0 = 0; k0 < C; k0 += 16)
              tc::mma_bf16(tmem_s, tc::plane_desc_mn(a_base, C, k0), tc::plane_desc_mn(b_base, C, k0), idesc, true);
          }
        }
        {
          constexpr uint32_t idesc = tc::idesc_bf16(128, C, /*a_mn=*/false, /*b_mn=*/true);
#pragma unroll
          for (int pass = 0; pass < 6; ++pass) {
            const uint32_t a_base = smem_u32(dAp + tc::bf16x3_term_a(pass) * Plan::DA_PLANE);
            const uint32_t b_base = smem_u32(Fp + tc::bf16x3_term_b(pass) * Plan::F_PLANE);
#pragma unroll
            for (int k0 = 0; k0 < NE; k0 += 16)
              tc::mma_bf16(tmem_dw, tc::plane_desc_k(a_base, C, k0), tc::plane_desc_mn(b_base, NE, k0), idesc,
                           !first_tile || (pass | k0) != 0);
          }
        }
        tc::mma_commit(&bars[1]);
      }
      ltc_wait(&bars[1], phase);
      tc::fence_after_sync();
      LTC_TS();

      // ---- E3: thread = channel m; encoder gradients stay in registers, x-gradients go through T
      if (ch_active) {
#pragma unroll 1
        for (int g = half * (TC / 2); g < (half + 1) * (TC / 2); ++g) {
          float d[KT];
#pragma unroll
          for (int k0 = 0; k0 < KT; k0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem_s + lane_base + (uint32_t)(g * KT + k0), v);
#pragma unroll
            for (int k = 0; k < 16; ++k) d[k0 + k] = v[k];
          }
          if (ch < H) {
            const int toff = ((ch >> 2) * (NE + 1) + g * KT) * 4 + (ch & 3);
#pragma unroll
            for (int k = 0; k < KT; ++k) T[toff + 4 * k] = d[k];
          } else {
            const int deg = DEG[g];
            const float4 p = P[g];
            const float sp = fmaf(mw0.z, p.z, fmaf(mw0.y, p.y, fmaf(mw0.x, p.x, mw1.w)));
            float gsum = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k)
              if (k < deg) {
                const float4 q = Q[g * KT + k];
                float s = sp;  // lrelu'(z) from the recomputed pre-activation
                s = fmaf(mw0.w, q.x, s), s = fmaf(mw1.x, q.y, s), s = fmaf(mw1.y, q.z, s);
                s = fmaf(mw1.z, q.w, s);
                const float dz = d[k] * (s > 0.f ? 1.f : kLReluSlope);
                gw[3] = fmaf(dz, q.x, gw[3]);
                gw[4] = fmaf(dz, q.y, gw[4]);
                gw[5] = fmaf(dz, q.z, gw[5]);
                gw[6] = fmaf(dz, q.w, gw[6]);
                gsum += dz;
              }
            gw[0] = fmaf(gsum, p.x, gw[0]);
            gw[1] = fmaf(gsum, p.y, gw[1]);
            gw[2] = fmaf(gsum, p.z, gw[2]);
            gw[7] += gsum;
          }
        }
      }
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      LTC_TS();
      // ---- scatter the x-gradients: thread = (edge, 4 channels), 16-byte vector reductions
      constexpr int H4 = H / 4;
#pragma unroll
      for (int t0 = 0; t0 < NE * H4; t0 += LTC_THREADS) {
        const int t = t0 + tid;
        if ((NE * H4) % LTC_THREADS == 0 || t < NE * H4) {
          const int e = t / H4, m4 = t % H4;
          const int j = NB[e];
          if (j >= 0) {
            const float4 v = *reinterpret_cast<const float4*>(T + (m4 * (NE + 1) + e) * 4);
            atomicAdd(reinterpret_cast<float4*>(grad_x + (int64_t)j * H) + m4, v);
          }
        }
      }
      first_tile = false;
      // the next tile's B1 only writes NB / Q / P / DEG, all of which the scatter above still reads
      __syncthreads();
      LTC_TS();
    }
    phase ^= 1u;
  }

  if constexpr (BWD) {
    if (!first_tile && ch_active) {
      // dW[n][m] of this CTA: thread = row n, the two warps of a lane quadrant split the columns
      for (int c0 = half * (C / 2); c0 < (half + 1) * (C / 2); c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_dw + lane_base + (uint32_t)c0, v);
        float4* dst = reinterpret_cast<float4*>(grad_att_w + (int64_t)ch * C + c0);
#pragma unroll
        for (int u = 0; u < 4; ++u) atomicAdd(dst + u, make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]));
      }
      if (ch >= H) {
#pragma unroll
        for (int t = 0; t < 7; ++t) atomicAdd(grad_enc_w + (ch - H) * 7 + t, gw[t]);
        atomicAdd(grad_enc_b + (ch - H), gw[7]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_s, Plan::TMEM_COLS);
  LTC_TS();
  if (dbg != nullptr && blockIdx.x == 0 && tid == 0) dbg[127] = nts;
#undef LTC_TS
}

template <int C, int NE, int KT, bool BWD, int MINB>
static int launch_lfa_tc(const float* x, const float* pos, const int32_t* nbr, const float* enc_w, const float* enc_b,
                         const float* att_w, float* out, const float* go, float* gx, float* gew, float* geb, float* gaw,
                         int64_t n, cudaStream_t st) {
  using Plan = LtcPlan<C, NE, KT, BWD>;
  auto kern = lfa_tc_kernel<C, NE, KT, BWD, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan::SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "lfa_tc smem attribute");
  const int64_t ntiles = ceil_div(n, Plan::TC);
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LTC_THREADS, Plan::SMEM_BYTES) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int by_tmem = 512 / (int)Plan::TMEM_COLS;  // co-resident CTAs must all get their TMEM columns
  if (per_sm > by_tmem) per_sm = by_tmem;
  int64_t grid = (int64_t)num_sms() * per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, LTC_THREADS, Plan::SMEM_BYTES, st>>>(x, pos, nbr, enc_w, enc_b, att_w, out, go, gx, gew, geb, gaw, n,
                                                             ntiles, tc_debug_buffer());
  B200_CHECK_LAUNCH("lfa_tc_kernel");
  return B200_OK;
}

// (C, KT, NE forward, NE backward, CTAs per SM the register allocator leaves room for)
#define B200_LFA_TC_CASES(X) \
  X(32, 16, 128, 64, 2) X(64, 16, 128, 64, 2) X(128, 16, 128, 64, 1) X(32, 32, 128, 64, 2) X(64, 32, 128, 64, 2) X(128, 32, 128, 64, 1)

// returns B200_E_UNSUPPORTED when this (c, kt) has no tensor-core kernel (the caller then uses the FMA kernel)
int lfa_tc_fwd_dispatch(const float* x, const float* pos, const int32_t* nbr, const float* enc_w, const float* enc_b,
                        const float* att_wt, float* out, int64_t n, int c, int kt, cudaStream_t st) {
  if (!tensor_cores_enabled()) return B200_E_UNSUPPORTED;
#define X(C_, KT_, NEF_, NEB_, MINB_)                                                                                     \
  if (c == C_ && kt == KT_)                                                                                               \
    return launch_lfa_tc<C_, NEF_, KT_, false, MINB_>(x, pos, nbr, enc_w, enc_b, att_wt, out, nullptr, nullptr, nullptr, \
                                                      nullptr, nullptr, n, st);
  B200_LFA_TC_CASES(X)
#undef X
  return B200_E_UNSUPPORTED;
}

int lfa_tc_bwd_dispatch(const float* x, const float* pos, const int32_t* nbr, const float* enc_w, const float* enc_b,
                        const float* att_w, const float* go, float* gx, float* gew, float* geb, float* gaw, int64_t n, int c,
                        int kt, cudaStream_t st) {
  if (!tensor_cores_enabled()) return B200_E_UNSUPPORTED;
#define X(C_, KT_, NEF_, NEB_, MINB_)                                                                                \
  if (c == C_ && kt == KT_)                                                                                          \
    return launch_lfa_tc<C_, NEB_, KT_, true, MINB_>(x, pos, nbr, enc_w, enc_b, att_w, nullptr, go, gx, gew, geb, gaw, n, st);
  B200_LFA_TC_CASES(X)
#undef X
  return B200_E_UNSUPPORTED;
}

bool lfa_tc_supported(int c, int kt) {
  if (!tensor_cores_enabled()) return false;
#define X(C_, KT_, NEF_, NEB_, MINB_) \
  if (c == C_ && kt == KT_) return true;
  B200_LFA_TC_CASES(X)
#undef X
  return false;
}

}  // namespace b200
