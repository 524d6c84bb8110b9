// Row-streaming NT GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32, TMEM accumulator):
//     out[i][m] = sum_k x[i][k] * wm[m][k] (+ bias[m])        x: [n, ktot] rows (two segments), wm: [mrows, ktot]
// = torch.nn.Linear forward of the SharedMLP layers (myria3d/models/modules/pyg_randla_net.py:97-109; wm = W) and
// its input gradient (x = grad_y, wm = W^T, transposed into a workspace by a tiny kernel first).
//
// Both operands are K-major in global memory already, so the loaders are plain 16-byte row reads (8 lanes cover one
// row's 128 contiguous bytes) -> tf32 hi/lo split -> STS.128 into the no-swizzle canonical layout of tc.cuh.
// Output CHANNELS sit on the 128 TMEM lanes and rows on the columns (like lfa_tc.cu): the epilogue thread owns one
// channel, so the bias add and the BatchNorm column statistics (sum, sum of squares in fp64, one partial per row
// tile, the format b200_bn_finalize consumes) are thread-local, and every store instruction of a warp writes 32
// consecutive channels of one row (128 contiguous bytes).
// CTA = 128 channels x BN rows; the K dimension streams through two shared-memory stages of 32; one thread issues
// 3 x 4 MMAs per stage and commits to the stage's mbarrier while the other threads already fetch the next stage.
#include "tc.cuh"

namespace b200 {

constexpr int TNT_THREADS = 256;
constexpr int TNT_KC = 32;

struct NtRows {  // [a1 | a2] rows; c1 % 32 == 0, every segment float4-addressable (checked by the caller)
  const float* a1;
  int64_t ld1;
  int c1;
  const float* a2;
  int64_t ld2;
  int c2;
};
struct NtOut {  // channel m < c1 -> o1[row * ld1 + m], else o2[row * ld2 + m - c1]; either pointer may be null
  float* o1;
  int64_t ld1;
  int c1;
  float* o2;
  int64_t ld2;
};

template <int BN>
__global__ void __launch_bounds__(TNT_THREADS, 1)
tc_nt_kernel(NtRows X, const float* __restrict__ wm, int mrows, const float* __restrict__ bias, NtOut O,
             double* __restrict__ colstats /* [row tiles][2 * mrows] or nullptr */, int64_t n, int m_tiles, int total_tiles,
             long long* __restrict__ dbg /* clock64 timeline of CTA 0 (scripts/nt_timeline.py); nullptr in production */) {
  extern __shared__ __align__(128) float tnt_smem[];
  const bool rec = dbg && blockIdx.x == 0 && threadIdx.x == 32;
  int nrec = 0;
#define B200_TS() do { if (rec && nrec < 120) dbg[nrec++] = clock64(); } while (0)
  B200_TS();  // [0]
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ double stat_sh[2][128][2];
  constexpr size_t W_FLOATS = tc::operand_floats(128, TNT_KC);
  constexpr size_t X_FLOATS = tc::operand_floats(BN, TNT_KC);
  constexpr size_t STAGE_FLOATS = 2 * W_FLOATS + 2 * X_FLOATS;
  constexpr int XQ = BN * 8 / TNT_THREADS;  // float4 of the row operand per thread and stage
  constexpr int WQ = 128 * 8 / TNT_THREADS;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int ktot = X.c1 + X.c2;
  const int nchunks = ktot / TNT_KC;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, 2 * BN);  // two accumulator buffers
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = tc::idesc_tf32(128, BN);
  bool alive = true;  // false after a barrier time-out: skip the remaining work, never hang

  // item = tid + q * 256 -> operand row item / 8, 16-byte k-group item % 8 (8 lanes read one row's 128 bytes;
  // their STS.128 hit 8 different bank groups because the k-group pitch (rows + 1) * 16 B is odd in 16-byte units)
  float4 xr[XQ], wr[WQ];
  auto load_chunk = [&](int tile, int k0) {  // tile -> (row tile, channel tile): neighbouring CTAs share the rows in L2
    const int m0 = (tile % m_tiles) * 128;
    const int64_t i0 = (int64_t)(tile / m_tiles) * BN;
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int item = tid + q * TNT_THREADS, r = item >> 3, k = k0 + 4 * (item & 7);
      const int64_t row = i0 + r;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n) {
        if (k < X.c1)
          t = __ldg(reinterpret_cast<const float4*>(X.a1 + row * X.ld1 + k));
        else
          t = __ldg(reinterpret_cast<const float4*>(X.a2 + row * X.ld2 + (k - X.c1)));
      }
      xr[q] = t;
    }
#pragma unroll
    for (int q = 0; q < WQ; ++q) {
      const int item = tid + q * TNT_THREADS, r = item >> 3, k = k0 + 4 * (item & 7);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < mrows) t = __ldg(reinterpret_cast<const float4*>(wm + (int64_t)(m0 + r) * ktot + k));
      wr[q] = t;
    }
  };
  auto split_store = [&](float* hi_base, float* lo_base, int rows_p1, int item, const float4& v) {
    const int off = ((item & 7) * rows_p1 + (item >> 3)) * 4;
    float4 hi, lo;
    tc::split_tf32(v.x, hi.x, lo.x), tc::split_tf32(v.y, hi.y, lo.y);
    tc::split_tf32(v.z, hi.z, lo.z), tc::split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    *reinterpret_cast<float4*>(lo_base + off) = lo;
  };

  // Epilogue of one tile out of TMEM buffer `buf`: thread = channel, TMEM columns = rows.  Two 16-column loads are kept
  // in flight (the next block travels while the current one is processed); the BatchNorm sums of a 16-row block are
  // taken in fp32 and accumulated across blocks in fp64.
  auto epilogue = [&](int tile, int buf) {
    const int m0 = (tile % m_tiles) * 128;
    const int64_t i0 = (int64_t)(tile / m_tiles) * BN;
    const int lane_q = warp & 3, half = warp >> 2;  // warp w reads TMEM lanes 32*(w%4) .. +31 and the row half w/4
    const int ml = lane_q * 32 + (tid & 31);
    const int m = m0 + ml;
    const bool mok = m < mrows;
    const float b = (bias && mok) ? __ldg(bias + m) : 0.f;
    float* obase = nullptr;
    int64_t old = 0;
    if (mok) {
      if (m < O.c1) {
        if (O.o1) obase = O.o1 + m, old = O.ld1;
      } else if (O.o2) {
        obase = O.o2 + (m - O.c1), old = O.ld2;
      }
    }
    double s1 = 0.0, s2 = 0.0;
    const bool st = obase != nullptr;
    float* const pbase = st ? obase : nullptr;
    const bool want_stats = colstats != nullptr;
    auto process = [&](const uint32_t (&r)[16], int64_t row0) {
      // (pointer bumps and predicated stores: a nullable pointer indexed with a 64-bit product per element compiled to a
      // divergent branch region + ~10 integer instructions per store and made the epilogue 4x longer than the MMAs)
      float y[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(r[j]) + b;
      const int64_t left = n - row0;
      const int lim = left >= 16 ? 16 : (left < 0 ? 0 : (int)left);
      float* p = pbase + row0 * old;
      if (lim == 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (st) *p = y[j];
          p += old;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (st && j < lim) *p = y[j];
          p += old;
        }
      }
      if (want_stats) {
        float sa = 0.f, sq = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = (j < lim) ? y[j] : 0.f;
          sa += v;
          sq = fmaf(v, v, sq);
        }
        s1 += (double)sa;
        s2 += (double)sq;
      }
    };
    const uint32_t tbase = tmem_d + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2));
    const int64_t rbase = i0 + half * (BN / 2);
    uint32_t ra[16], rb[16];
    tc::tmem_ld16_async(tbase, ra);
    tc::tmem_ld_wait(ra);
#pragma unroll 1
    for (int cc = 0; cc < BN / 2; cc += 32) {
      tc::tmem_ld16_async(tbase + (uint32_t)(cc + 16), rb);
      process(ra, rbase + cc);
      tc::tmem_ld_wait(rb);
      if (cc + 32 < BN / 2) tc::tmem_ld16_async(tbase + (uint32_t)(cc + 32), ra);
      process(rb, rbase + cc + 16);
      if (cc + 32 < BN / 2) tc::tmem_ld_wait(ra);
    }
    if (colstats) {
      stat_sh[half][ml][0] = s1;
      stat_sh[half][ml][1] = s2;
    }
    B200_TS();  // tile: epilogue done (this warp)
    tc::fence_before_sync();
    __syncthreads();  // this TMEM buffer is drained (a later tile's first MMA overwrites it), stat_sh complete
    tc::fence_after_sync();
    if (colstats && tid < 128 && m0 + tid < mrows) {
      double* part = colstats + (int64_t)(tile / m_tiles) * 2 * mrows;
      part[m0 + tid] = stat_sh[0][tid][0] + stat_sh[1][tid][0];
      part[mrows + m0 + tid] = stat_sh[0][tid][1] + stat_sh[1][tid][1];
    }
  };

  // Persistent CTA: tiles blockIdx.x, + gridDim.x, ...  `g` counts the K-chunks this CTA has issued over all its tiles:
  // chunk g uses stage g & 1, whose mbarrier completes its (g >> 1)-th phase when the chunk's MMAs are done.
  // Tile t accumulates into TMEM buffer t & 1; its epilogue runs AFTER the MMAs of tile t+1 have been issued into the
  // other buffer, so the tensor pipe works while the previous tile is read out and stored, and the first chunk of the
  // tile after that is already travelling from DRAM.
  int g = 0, it = 0, prev_tile = -1;
  B200_TS();  // [1] after alloc/init
  if ((int)blockIdx.x < total_tiles) load_chunk(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total_tiles && alive; tile += gridDim.x, ++it) {
    const uint32_t tmem_acc = tmem_d + (uint32_t)((it & 1) * BN);
    for (int s = 0; s < nchunks && alive; ++s, ++g) {
      const int stage = g & 1;
      float* Wh = tnt_smem + stage * STAGE_FLOATS;
      float* Wl = Wh + W_FLOATS;
      float* Xh = Wl + W_FLOATS;
      float* Xl = Xh + X_FLOATS;
      // the stage may be overwritten only after the MMAs of chunk g-2 (its previous user) have completed
      if (g >= 2) {
        alive = tc::mbar_wait_bounded(&bars[stage], (uint32_t)(((g - 2) >> 1) & 1));
        alive = __syncthreads_and(alive) != 0;
      }
      if (!alive) break;
      B200_TS();  // chunk: stage free
#pragma unroll
      for (int q = 0; q < WQ; ++q) split_store(Wh, Wl, 129, tid + q * TNT_THREADS, wr[q]);
#pragma unroll
      for (int q = 0; q < XQ; ++q) split_store(Xh, Xl, BN + 1, tid + q * TNT_THREADS, xr[q]);
      if (s + 1 < nchunks)
        load_chunk(tile, (s + 1) * TNT_KC);  // prefetch: consumed next iteration
      else if (tile + (int)gridDim.x < total_tiles)
        load_chunk(tile + gridDim.x, 0);     // ... or by the next tile
      B200_TS();  // chunk: stored + next loads issued

      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      tc::fence_after_sync();
      B200_TS();  // chunk: after barrier
      if (tid == 0) {
        const uint32_t lbo_w = tc::lbo_bytes(128), lbo_x = tc::lbo_bytes(BN);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_base = smem_u32((pass == 1) ? Wl : Wh);  // hi*hi, lo*hi, hi*lo
          const uint32_t b_base = smem_u32((pass == 2) ? Xl : Xh);
#pragma unroll
          for (int ks = 0; ks < TNT_KC / 8; ++ks) {
            const uint64_t ad = tc::smem_desc(a_base + (uint32_t)(2 * ks) * lbo_w, lbo_w, tc::kSboBytes);
            const uint64_t bd = tc::smem_desc(b_base + (uint32_t)(2 * ks) * lbo_x, lbo_x, tc::kSboBytes);
            tc::mma_tf32(tmem_acc, ad, bd, idesc, (s | pass | ks) != 0);
          }
        }
        tc::mma_commit(&bars[stage]);
      }
    }
    if (!alive) break;
    // the previous tile: with >= 2 chunks per tile its last chunk (g - nchunks - 1) was already awaited by this tile's
    // second chunk (stage reuse); with one chunk per tile nothing has been committed to its barrier since, so the
    // parity wait below is unambiguous
    if (prev_tile >= 0) {
      if (nchunks == 1) {
        const int last = g - 2;
        alive = tc::mbar_wait_bounded(&bars[last & 1], (uint32_t)((last >> 1) & 1));
        alive = __syncthreads_and(alive) != 0;
        tc::fence_after_sync();
      }
      B200_TS();  // tile: previous tile's MMAs complete
      if (alive) epilogue(prev_tile, (it - 1) & 1);
    }
    prev_tile = tile;
  }
  if (alive && prev_tile >= 0) {  // the last tile of this CTA
    const int last = g - 1;
    alive = tc::mbar_wait_bounded(&bars[last & 1], (uint32_t)((last >> 1) & 1));
    alive = __syncthreads_and(alive) != 0;
    tc::fence_after_sync();
    B200_TS();
    if (alive) epilogue(prev_tile, (it - 1) & 1);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, 2 * BN);
  if (rec) dbg[127] = nrec;
#undef B200_TS
}

long long* tc_debug_buffer();  // tc_gemm.cu

// wt[k][m] = w[m][k]   (w: [rows, cols] row-major)
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (r0 + r < rows && c0 + tx < cols) tile[r][tx] = __ldg(w + (int64_t)(r0 + r) * cols + c0 + tx);
  __syncthreads();
  for (int c = ty; c < 32; c += 8)
    if (c0 + c < cols && r0 + tx < rows) wt[(int64_t)(c0 + c) * rows + r0 + tx] = tile[tx][c];
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// sizes only (the statistics-partial count must be known before the launch): see b200_linear_fwd_num_stat_partials
bool tc_nt_shape_ok(int64_t n, int c1, int c2, int mrows) {
  const int ktot = c1 + c2;
  return tensor_cores_enabled() && n >= 512 && mrows >= 64 && ktot >= 32 && ktot % TNT_KC == 0 && c1 % TNT_KC == 0;
}
int tc_nt_rows_per_tile(int64_t n, int mrows) {
  // 128-row tiles only while they all fit one wave of CTAs; 64-row tiles when 128-row ones would leave more than
  // half of the SMs idle (levels 3-4: 3 200 / 800 rows)
  const int64_t tiles128 = ceil_div(n, 128) * ceil_div(mrows, 128);
  if (tiles128 > num_sms()) return 256;
  return (2 * tiles128 <= num_sms()) ? 64 : 128;
}

template <int BN>
static int launch_tc_nt_bn(const NtRows& X, const float* wm, int mrows, const float* bias, const NtOut& O, double* colstats,
                           int64_t n, cudaStream_t st) {
  const size_t smem = sizeof(float) * 2 * (2 * tc::operand_floats(128, TNT_KC) + 2 * tc::operand_floats(BN, TNT_KC));
  auto kern = tc_nt_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "tc_nt smem attribute");
  const int m_tiles = (int)ceil_div(mrows, 128);
  const int64_t total = (int64_t)m_tiles * ceil_div(n, BN);
  B200_REQUIRE(total < (1ll << 30), B200_E_UNSUPPORTED, "tensor-core linear layer: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());  // persistent: one CTA per SM
  kern<<<grid, TNT_THREADS, smem, st>>>(X, wm, mrows, bias, O, colstats, n, m_tiles, (int)total, tc_debug_buffer());
  B200_CHECK_LAUNCH("tc_nt_kernel");
  return B200_OK;
}

// out[i][m] = sum_k [a1|a2][i][k] * wm[m][k] + bias[m]; out split into two channel segments (o2 may be null / oc1 = mrows)
int launch_tc_nt(const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2, const float* wm, int mrows,
                 const float* bias, float* o1, int64_t old1, int oc1, float* o2, int64_t old2, double* colstats, int64_t n,
                 cudaStream_t st) {
  B200_REQUIRE(al16(a1) && ld1 % 4 == 0 && (c2 == 0 || (al16(a2) && ld2 % 4 == 0)) && al16(wm), B200_E_INVALID,
               "tensor-core linear layer: activation rows and weights must be 16-byte aligned (row strides %% 4 == 0)");
  const NtRows X{a1, ld1, c1, a2, ld2, c2};
  const NtOut O{o1, old1, oc1, o2, old2};
  const int rows = tc_nt_rows_per_tile(n, mrows);
  if (rows == 256) return launch_tc_nt_bn<256>(X, wm, mrows, bias, O, colstats, n, st);
  if (rows == 64) return launch_tc_nt_bn<64>(X, wm, mrows, bias, O, colstats, n, st);
  return launch_tc_nt_bn<128>(X, wm, mrows, bias, O, colstats, n, st);
}

int launch_transpose(const float* w, float* wt, int rows, int cols, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  transpose_kernel<<<grid, 256, 0, st>>>(w, wt, rows, cols);
  B200_CHECK_LAUNCH("transpose_kernel");
  return B200_OK;
}

}  // namespace b200
