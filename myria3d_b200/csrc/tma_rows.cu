// Row-streaming Linear kernels of levels 0-1 fed by TENSOR TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) for sm_100a.
//
// Replaces, for the narrow layers on >= 8192 rows (16/32/64 channels on both sides), the three GEMMs of
// torch.nn.Linear inside SharedMLP / FPModule (myria3d/models/modules/pyg_randla_net.py:97-109, :251, :42-53):
//   NN  y[n, M]  = [a1 | a2][n, K] . B[K, M] + bias      forward (B = W^T) and input gradient (rows = grad_y, B = W)
//   TN  gw[M, K] += gy[n, M]^T . [a1 | a2][n, K],  gb += colsum(gy)     weight / bias gradient
// Those layers move 13-80 MB for 0.03-0.8 GFLOP: they are HBM-bound, and the register-staged tile GEMMs of pointwise.cu
// sit at 1.0-2.0 TB/s because every CTA exposes one full DRAM latency per 16-column slice.  Here:
//   * persistent CTAs (2 per SM), each walking the 128-row tiles blockIdx.x, blockIdx.x + gridDim.x, ...;
//   * the row tiles travel global -> shared memory as 2-D TMA boxes (one elected thread issues, no registers, no
//     address arithmetic in the other 255 threads) through a 3-4 stage ring of mbarriers (expect_tx / try_wait.parity):
//     2 CTAs x 3-4 stages x 8-32 KB in flight per SM covers the bandwidth-delay product of HBM3e (about 44 KB per SM);
//   * the boxes are at most 128 bytes wide and use the hardware XOR swizzle of their width (SWIZZLE_128B / 64B), so the
//     float4 operand reads of the FMA loop are bank-conflict free without padding;
//   * weights sit in shared memory for the whole kernel; accumulators in registers (4 x TR outputs per thread, or a
//     4 x 4 block of gw); BatchNorm statistics (fp64 sum, sum of squares) are carried in registers across ALL tiles of
//     the CTA and reduced once at the end: one partial row per CTA instead of one per 128-row tile.
// The tensor map (CUtensorMap) is encoded on the host for every call (cuTensorMapEncodeTiled through
// cudaGetDriverEntryPoint: the library keeps loading on a box without libcuda) and passed as a __grid_constant__
// kernel parameter, so a captured CUDA graph carries it inside the kernel node.
#include <cuda.h>

#include "common.cuh"

namespace b200 {

namespace {

constexpr int TR_ROWS = 128;  // rows per tile (the TMA box height)
constexpr int TR_THREADS = 256;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// rows x cols fp32 matrix with row stride ld (floats); box = TR_ROWS rows x ws columns, swizzle of the box width
bool make_map(CUtensorMap* m, const float* base, int64_t ld, int cols, int64_t n, int ws, int rows = TR_ROWS) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)n};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)ws, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = ws == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int col, int row, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(bar)
      : "memory");
}

// byte offset of the 16-byte chunk `chunk` of row `row` inside a box of WS floats per row (tile base 1024-byte
// aligned): the TMA swizzle XORs address bits [4, 4 + log2(WS / 8)) with bits [7, ...) -- Swizzle<3,4,3> for 128-byte
// rows, Swizzle<2,4,3> for 64-byte rows
template <int WS>
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  const uint32_t l = (uint32_t)row * (WS * 4) + (uint32_t)chunk * 16;
  constexpr uint32_t MASK = WS == 32 ? 7u : 3u;
  return l ^ (((l >> 7) & MASK) << 4);
}

template <int WS, int NSUB, int MP>
struct NnCfg {
  static constexpr int KP = WS * NSUB;
  static constexpr int SUB_BYTES = TR_ROWS * WS * 4;
  static constexpr int STAGE_BYTES = NSUB * SUB_BYTES;
  static constexpr int STAGES = STAGE_BYTES >= 32 * 1024 ? 3 : 4;
  static constexpr int NCG = MP / 4;               // column groups of 4 outputs
  static constexpr int NRG = TR_THREADS / NCG;     // row groups
  static constexpr int TR = TR_ROWS / NRG;         // rows per thread: row = i * NRG + rg
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + (size_t)KP * MP * 4 + STAGES * 8;
  static_assert(TR >= 1 && NCG <= 32, "thread tile");
};

// ------------------------------------------------------------------ y = [x0 | x1] B + bias (+ BatchNorm partial sums)
template <int WS, int NSUB, int MP>
__global__ void __launch_bounds__(TR_THREADS, 2)
tma_rows_nn_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, int col1,
                   const float* __restrict__ w, int w_ld, int w_out_major, const float* __restrict__ bias,
                   float* __restrict__ o1, int64_t old1, int oc1, float* __restrict__ o2, int64_t old2, int64_t n,
                   int ntiles, double* __restrict__ colstats, int num_partials) {
  using C = NnCfg<WS, NSUB, MP>;
  extern __shared__ uint8_t tr_smem_raw[];
  const uint32_t raw = smem_u32(tr_smem_raw);
  uint8_t* base = tr_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  float* Bs = reinterpret_cast<float*>(base + C::STAGES * C::STAGE_BYTES);  // [KP][MP]
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + C::KP * MP);
  const int tid = threadIdx.x;
  const int cg = tid % C::NCG, rg = tid / C::NCG;

  // The weights are fetched FIRST, all loads of a thread in flight at once (ncu, round 2: a load -> store loop here made
  // four dependent L2 round trips and held 20 % of the kernel's stall samples), and parked in registers while the
  // barriers are set up and the first tiles are requested.
  constexpr int NWREG = C::KP * MP / TR_THREADS;
  static_assert(C::KP * MP % TR_THREADS == 0, "weight staging");
  float wreg[NWREG];
#pragma unroll
  for (int i = 0; i < NWREG; ++i) {
    const int t = tid + i * TR_THREADS;  // global reads coalesced in both orientations
    wreg[i] = w_out_major ? __ldg(w + (int64_t)(t / C::KP) * w_ld + (t % C::KP)) : __ldg(w + (int64_t)(t / MP) * w_ld + (t % MP));
  }
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const CUtensorMap* const pm0 = &map0;
  const CUtensorMap* const pm1 = &map1;
  auto issue = [&, pm0, pm1](int i) {  // one thread: arm the stage's barrier with the byte count, then the box copies
    const int s = i % C::STAGES;
    const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * TR_ROWS;
    const uint32_t bar = smem_u32(&full[s]);
    const uint32_t dst = smem_u32(base + s * C::STAGE_BYTES);
    mbar_expect_tx(&full[s], C::STAGE_BYTES);
    tma_load_2d(dst, pm0, 0, row0, bar);
    if constexpr (NSUB == 2) tma_load_2d(dst + C::SUB_BYTES, pm1, col1, row0, bar);
  };
  if (tid == 0)
    for (int i = 0; i < C::STAGES && i < my_tiles; ++i) issue(i);
#pragma unroll
  for (int i = 0; i < NWREG; ++i) {  // Bs[k][m]
    const int t = tid + i * TR_THREADS;
    Bs[w_out_major ? (t % C::KP) * MP + (t / C::KP) : t] = wreg[i];
  }
  __syncthreads();

  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int u = 0; u < 4; ++u) bv[u] = __ldg(bias + cg * 4 + u);
  }
  // BatchNorm statistics in fp64 (exact products, no cancellation in E[y^2] - E[y]^2), carried across all tiles of the CTA.
  // (A pivoted fp32 variant -- sums of v - first value, fp64 only at the end -- was measured: same time, 25.6 vs 24.6 us.)
  double ps[4] = {0.0, 0.0, 0.0, 0.0}, pq[4] = {0.0, 0.0, 0.0, 0.0};
  const int col = cg * 4;
  float* const obase = (col < oc1) ? (o1 ? o1 + col : nullptr) : (o2 ? o2 + (col - oc1) : nullptr);
  const int64_t old = (col < oc1) ? old1 : old2;

  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % C::STAGES;
    mbar_wait(&full[s], (uint32_t)(i / C::STAGES) & 1u);
    float acc[C::TR][4];
#pragma unroll
    for (int r = 0; r < C::TR; ++r)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[r][u] = 0.f;
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const uint8_t* T = base + s * C::STAGE_BYTES + sub * C::SUB_BYTES;
#pragma unroll
      for (int j = 0; j < WS / 4; ++j) {
        float4 b[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(Bs + (sub * WS + j * 4 + kk) * MP + col);
#pragma unroll
        for (int r = 0; r < C::TR; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(T + sw_off<WS>(r * C::NRG + rg, j));
          // packed FMAs (FFMA2) over column pairs, the row value in the broadcast slot: half the issue slots and half the
          // FMA-pipe cycles of 16 scalar FFMAs (the scalar form issues every second cycle on sm_100a)
          ffma2_bc(a.x, b[0].x, b[0].y, acc[r][0], acc[r][1]), ffma2_bc(a.x, b[0].z, b[0].w, acc[r][2], acc[r][3]);
          ffma2_bc(a.y, b[1].x, b[1].y, acc[r][0], acc[r][1]), ffma2_bc(a.y, b[1].z, b[1].w, acc[r][2], acc[r][3]);
          ffma2_bc(a.z, b[2].x, b[2].y, acc[r][0], acc[r][1]), ffma2_bc(a.z, b[2].z, b[2].w, acc[r][2], acc[r][3]);
          ffma2_bc(a.w, b[3].x, b[3].y, acc[r][0], acc[r][1]), ffma2_bc(a.w, b[3].z, b[3].w, acc[r][2], acc[r][3]);
        }
      }
    }
    __syncthreads();  // every thread has read stage s: it can be refilled
    if (tid == 0 && i + C::STAGES < my_tiles) issue(i + C::STAGES);

    const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TR_ROWS;
#pragma unroll
    for (int r = 0; r < C::TR; ++r) {
      const int64_t row = row0 + r * C::NRG + rg;
      if (row >= n) continue;
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = acc[r][u] + bv[u];
      if (obase) *reinterpret_cast<float4*>(obase + row * old) = make_float4(v[0], v[1], v[2], v[3]);
      if (colstats) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ps[u] += (double)v[u];
          pq[u] = fma((double)v[u], (double)v[u], pq[u]);
        }
      }
    }
  }

  if (colstats) {
    // lanes of a warp with the same column group differ in lane bits >= log2(NCG)
#pragma unroll
    for (int off = C::NCG; off < 32; off <<= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ps[u] += __shfl_xor_sync(0xffffffffu, ps[u], off);
        pq[u] += __shfl_xor_sync(0xffffffffu, pq[u], off);
      }
    }
    __syncthreads();  // all tiles consumed, no copy in flight: the ring is free
    double* red = reinterpret_cast<double*>(base);  // [8 warps][2][MP]
    const int warp = tid >> 5, lane = tid & 31;
    if (lane < C::NCG) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        red[(warp * 2 + 0) * MP + col + u] = ps[u];
        red[(warp * 2 + 1) * MP + col + u] = pq[u];
      }
    }
    __syncthreads();
    if (tid < 2 * MP) {
      const int kind = tid / MP, c = tid % MP;
      double a = 0.0;
#pragma unroll
      for (int wq = 0; wq < TR_THREADS / 32; ++wq) a += red[(wq * 2 + kind) * MP + c];
      colstats[(int64_t)blockIdx.x * 2 * MP + kind * MP + c] = a;
      // the caller's buffer has one row per tile of the kernel this one stands in for: the rows beyond the grid are zero
      for (int64_t p = (int64_t)gridDim.x + blockIdx.x; p < num_partials; p += gridDim.x) colstats[p * 2 * MP + kind * MP + c] = 0.0;
    }
  }
}

// ------------------------------------------------------------------ gw += gy^T [x0 | x1], gb += colsum(gy)
// X rows: NS0 boxes of WS columns from map 0 (columns 0, WS, 2 WS, ...) followed by NS1 boxes from map 1 (the second
// segment of a concatenated input, or nothing); gy rows: MP columns in boxes of <= 32.  ROWS = box height (64 for the
// wide shapes, so that three to four stages fit).  A thread owns a 4 (m) x 4 KPT (k) block of gw for one slice of rows.
template <int WS, int NS0, int NS1, int MP, int ROWS>
struct TnCfg {
  static constexpr int NSUB = NS0 + NS1;
  static constexpr int KP = WS * NSUB;
  static constexpr int GWS = MP >= 32 ? 32 : MP;  // box width of the gy tile
  static constexpr int MSUB = MP / GWS;
  static constexpr int XSUB_BYTES = ROWS * WS * 4;
  static constexpr int GSUB_BYTES = ROWS * GWS * 4;
  static constexpr int STAGE_BYTES = NSUB * XSUB_BYTES + MSUB * GSUB_BYTES;
  static constexpr int STAGES = STAGE_BYTES >= 32 * 1024 ? 3 : 4;
  static constexpr int MINB = STAGE_BYTES > 32 * 1024 ? 1 : 2;
  static constexpr int NKB = KP / 4, NMB = MP / 4;
  static constexpr int KPT = (NKB * NMB > TR_THREADS) ? (NKB * NMB / TR_THREADS) : 1;  // 4-column groups per thread
  static constexpr int NKG = NKB / KPT;                                              // k groups (of 4 KPT columns)
  static constexpr int NBLK = NKG * NMB;
  static constexpr int RQ = TR_THREADS / NBLK;  // row slices per tile
  static constexpr int RPT = ROWS / RQ;         // rows per thread and tile
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + (size_t)(MP * KP + MP) * 4 + STAGES * 8;
  static_assert(NBLK <= TR_THREADS && RQ >= 1 && RQ * NBLK == TR_THREADS && RPT >= 1 && RPT * RQ == ROWS, "block map");
  static_assert(KPT == 1 || KPT == 2, "a thread owns 4 or 8 columns of gw");
  static_assert(KPT == 1 || (WS / 4) % KPT == 0, "a thread's columns stay inside one box");
  static_assert(XSUB_BYTES % 1024 == 0 && GSUB_BYTES % 1024 == 0, "swizzle atoms");
  static_assert(SMEM <= 232448, "shared memory");
};

template <int WS, int NS0, int NS1, int MP, int ROWS>
__global__ void __launch_bounds__(TR_THREADS, TnCfg<WS, NS0, NS1, MP, ROWS>::MINB)
tma_rows_tn_kernel(const __grid_constant__ CUtensorMap mapx0, const __grid_constant__ CUtensorMap mapx1,
                   const __grid_constant__ CUtensorMap mapg, float* __restrict__ gw, int gw_ld, float* __restrict__ gb,
                   int ntiles) {
  using C = TnCfg<WS, NS0, NS1, MP, ROWS>;
  extern __shared__ uint8_t tr_smem_raw[];
  const uint32_t raw = smem_u32(tr_smem_raw);
  uint8_t* base = tr_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  float* R = reinterpret_cast<float*>(base + C::STAGES * C::STAGE_BYTES);  // [MP][KP] + [MP]
  uint64_t* full = reinterpret_cast<uint64_t*>(R + MP * C::KP + MP);
  const int tid = threadIdx.x;
  const int bid = tid % C::NBLK, rq = tid / C::NBLK;
  const int kg = bid % C::NKG, mb = bid / C::NKG;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  for (int t = tid; t < MP * C::KP + MP; t += TR_THREADS) R[t] = 0.f;
  __syncthreads();

  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const CUtensorMap* const pm0 = &mapx0;
  const CUtensorMap* const pm1 = &mapx1;
  const CUtensorMap* const pmg = &mapg;
  auto issue = [&, pm0, pm1, pmg](int i) {
    const int s = i % C::STAGES;
    const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * ROWS;
    const uint32_t bar = smem_u32(&full[s]);
    uint32_t dst = smem_u32(base + s * C::STAGE_BYTES);
    mbar_expect_tx(&full[s], C::STAGE_BYTES);
#pragma unroll
    for (int q = 0; q < NS0; ++q) tma_load_2d(dst + q * C::XSUB_BYTES, pm0, q * WS, row0, bar);
#pragma unroll
    for (int q = 0; q < NS1; ++q) tma_load_2d(dst + (NS0 + q) * C::XSUB_BYTES, pm1, q * WS, row0, bar);
    dst += C::NSUB * C::XSUB_BYTES;
#pragma unroll
    for (int ms = 0; ms < C::MSUB; ++ms) tma_load_2d(dst + ms * C::GSUB_BYTES, pmg, ms * C::GWS, row0, bar);
  };
  if (tid == 0)
    for (int i = 0; i < C::STAGES && i < my_tiles; ++i) issue(i);

  float acc[C::KPT][4][4];  // acc[j][m][k]
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < C::KPT; ++j)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[j][a][b] = 0.f;
  // which box / 16-byte chunk of a row this thread reads (fixed for the whole kernel)
  const int k0 = kg * 4 * C::KPT;  // first gw column of the thread
  const int xsub = k0 / WS, xchunk = (k0 % WS) / 4;
  const int gsub = (mb * 4) / C::GWS, gchunk = ((mb * 4) % C::GWS) / 4;

  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % C::STAGES;
    mbar_wait(&full[s], (uint32_t)(i / C::STAGES) & 1u);
    const uint8_t* X = base + s * C::STAGE_BYTES + xsub * C::XSUB_BYTES;
    const uint8_t* G = base + s * C::STAGE_BYTES + C::NSUB * C::XSUB_BYTES + gsub * C::GSUB_BYTES;
    // rows past n are zero-filled by the TMA unit: they add nothing
#pragma unroll 8
    for (int rr = 0; rr < C::RPT; ++rr) {
      const int row = rq * C::RPT + rr;
      const float4 g = *reinterpret_cast<const float4*>(G + sw_off<C::GWS>(row, gchunk));
#pragma unroll
      for (int j = 0; j < C::KPT; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(X + sw_off<WS>(row, xchunk + j));
        ffma2_bc(g.x, x.x, x.y, acc[j][0][0], acc[j][0][1]), ffma2_bc(g.x, x.z, x.w, acc[j][0][2], acc[j][0][3]);
        ffma2_bc(g.y, x.x, x.y, acc[j][1][0], acc[j][1][1]), ffma2_bc(g.y, x.z, x.w, acc[j][1][2], acc[j][1][3]);
        ffma2_bc(g.z, x.x, x.y, acc[j][2][0], acc[j][2][1]), ffma2_bc(g.z, x.z, x.w, acc[j][2][2], acc[j][2][3]);
        ffma2_bc(g.w, x.x, x.y, acc[j][3][0], acc[j][3][1]), ffma2_bc(g.w, x.z, x.w, acc[j][3][2], acc[j][3][3]);
      }
      if (kg == 0) bs[0] += g.x, bs[1] += g.y, bs[2] += g.z, bs[3] += g.w;
    }
    __syncthreads();
    if (tid == 0 && i + C::STAGES < my_tiles) issue(i + C::STAGES);
  }

  // CTA reduction over the row slices in shared memory, then one set of global reductions per CTA
#pragma unroll
  for (int j = 0; j < C::KPT; ++j)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if constexpr (C::RQ > 1)
          atomicAdd(&R[(mb * 4 + a) * C::KP + k0 + 4 * j + b], acc[j][a][b]);
        else
          R[(mb * 4 + a) * C::KP + k0 + 4 * j + b] = acc[j][a][b];  // one writer per element
      }
  if (kg == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) atomicAdd(&R[MP * C::KP + mb * 4 + a], bs[a]);
  }
  __syncthreads();
  for (int t = tid; t < MP * C::KP; t += TR_THREADS) atomicAdd(gw + (int64_t)(t / C::KP) * gw_ld + (t % C::KP), R[t]);
  if (gb)
    for (int t = tid; t < MP; t += TR_THREADS) atomicAdd(gb + t, R[MP * C::KP + t]);
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// the K side of a call: one segment of 16 / 32 / 64 columns or two segments of 32
struct KSide {
  int ws, nsub;
  bool ok;
};
KSide k_side(const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2) {
  KSide k{0, 0, false};
  if (!a1 || !al16(a1) || ld1 % 4 != 0 || ld1 < c1) return k;
  if (c2 == 0) {
    if (c1 == 16) k = {16, 1, true};
    else if (c1 == 32) k = {32, 1, true};
    else if (c1 == 64) k = {32, 2, true};
  } else if (c1 == 32 && c2 == 32 && a2 && al16(a2) && ld2 % 4 == 0 && ld2 >= c2) {
    k = {32, 2, true};
  }
  return k;
}

template <int WS, int NSUB, int MP>
int nn_go(const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2,
          const float* w, int w_ld, bool w_out_major, const float* bias, float* o1, int64_t old1, int oc1, float* o2,
          int64_t old2, int64_t n, double* colstats, int num_partials, cudaStream_t st) {
  using C = NnCfg<WS, NSUB, MP>;
  auto kern = tma_rows_nn_kernel<WS, NSUB, MP>;
  const int ntiles = (int)ceil_div(n, TR_ROWS);
  static thread_local int occ = -1;
  if (occ < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    int o = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, TR_THREADS, C::SMEM);
    if (e != cudaSuccess || o < 1) return cuda_fail(e == cudaSuccess ? cudaErrorLaunchOutOfResources : e, "tma_rows_nn occupancy");
    occ = o > 2 ? 2 : o;
  }
  const int64_t cap = (int64_t)num_sms() * occ;
  int grid = (int)(ntiles < cap ? ntiles : cap);
  if (colstats && grid > num_partials) grid = num_partials;  // one partial row per CTA, the caller sized the buffer
  CUtensorMap m0, m1;
  if (!make_map(&m0, a1, ld1, c1, n, WS)) {
    set_error("tma_rows: cuTensorMapEncodeTiled failed (a1 %p ld %lld c %d n %lld)", (const void*)a1, (long long)ld1, c1, (long long)n);
    return B200_E_CUDA;
  }
  int col1 = 0;
  m1 = m0;
  if (NSUB == 2) {
    if (c2 > 0) {
      if (!make_map(&m1, a2, ld2, c2, n, WS)) {
        set_error("tma_rows: cuTensorMapEncodeTiled failed (second segment)");
        return B200_E_CUDA;
      }
    } else {
      col1 = WS;
    }
  }
  kern<<<grid, TR_THREADS, C::SMEM, st>>>(m0, m1, col1, w, w_ld, w_out_major ? 1 : 0, bias, o1, old1, oc1, o2, old2, n, ntiles,
                                         colstats, num_partials);
  B200_CHECK_LAUNCH("tma_rows_nn_kernel");
  return B200_OK;
}

// one TN launch: X = NS0 boxes of a1 (+ NS1 boxes of a2), gw columns [0, WS (NS0 + NS1)) of a row-major matrix with leading
// dimension gw_ld (the caller offsets gw when the weight's columns are covered by two launches)
template <int WS, int NS0, int NS1, int MP, int ROWS>
int tn_go(const float* gy, const float* a1, int64_t ld1, const float* a2, int64_t ld2, float* gw, int gw_ld, float* gb,
          int64_t n, cudaStream_t st) {
  using C = TnCfg<WS, NS0, NS1, MP, ROWS>;
  auto kern = tma_rows_tn_kernel<WS, NS0, NS1, MP, ROWS>;
  const int ntiles = (int)ceil_div(n, ROWS);
  static thread_local int occ = -1;
  if (occ < 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    int o = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, TR_THREADS, C::SMEM);
    if (e != cudaSuccess || o < 1) return cuda_fail(e == cudaSuccess ? cudaErrorLaunchOutOfResources : e, "tma_rows_tn occupancy");
    occ = o > 2 ? 2 : o;
  }
  const int64_t cap = (int64_t)num_sms() * occ;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  CUtensorMap m0, m1, mg;
  if (!make_map(&m0, a1, ld1, WS * NS0, n, WS, ROWS) || !make_map(&mg, gy, MP, MP, n, C::GWS, ROWS)) {
    set_error("tma_rows: cuTensorMapEncodeTiled failed (weight gradient)");
    return B200_E_CUDA;
  }
  m1 = m0;
  if (NS1 > 0 && !make_map(&m1, a2, ld2, WS * NS1, n, WS, ROWS)) {
    set_error("tma_rows: cuTensorMapEncodeTiled failed (second segment)");
    return B200_E_CUDA;
  }
  kern<<<grid, TR_THREADS, C::SMEM, st>>>(m0, m1, mg, gw, gw_ld, gb, ntiles);
  B200_CHECK_LAUNCH("tma_rows_tn_kernel");
  return B200_OK;
}

bool m_ok(int m) { return m == 16 || m == 32 || m == 64; }
bool seg_ok(const float* a, int64_t ld, int c) { return a && al16(a) && ld % 4 == 0 && ld >= c; }

}  // namespace

// ---- dispatch (declared in common.cuh)
// forward / input gradient: rows [a1 | a2] (K side), `mcols` outputs split over (o1: oc1 columns, o2: the rest)
bool tma_rows_nn_ok(int64_t n, const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2, int mcols,
                    const float* o1, int64_t old1, int oc1, const float* o2, int64_t old2) {
  if (n < 8192 || n > (int64_t)1 << 30 || !m_ok(mcols) || !encode_fn()) return false;
  if (!k_side(a1, ld1, c1, a2, ld2, c2).ok) return false;
  // 64 x 64 (one CTA per SM, 12 LDS.128 per 128 FFMA): the tcgen05 forward (tc_nt.cu) and the 64 x 64-tile FMA input
  // gradient are faster (22.5 vs 26.7 / 28.7 us on 51 200 rows); every other shape is faster here (profiles/)
  if (c1 + c2 == 64 && mcols == 64) return false;
  if (oc1 % 4 != 0 || oc1 < 0 || oc1 > mcols) return false;
  if (o1 && (!al16(o1) || old1 % 4 != 0)) return false;
  if (o2 && (!al16(o2) || old2 % 4 != 0)) return false;
  return true;
}

int launch_tma_rows_nn(const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2,
                       const float* w, int w_ld, bool w_out_major, const float* bias, float* o1, int64_t old1, int oc1,
                       float* o2, int64_t old2, int mcols, int64_t n, double* colstats, int num_partials, cudaStream_t st) {
  if (colstats && num_partials < 1) {
    set_error("tma_rows: BatchNorm statistics need at least one partial row");
    return B200_E_INVALID;
  }
  const KSide k = k_side(a1, ld1, c1, a2, ld2, c2);
#define B200_NN(WS_, NS_, MP_)                                                                                              \
  if (k.ws == WS_ && k.nsub == NS_ && mcols == MP_)                                                                         \
    return nn_go<WS_, NS_, MP_>(a1, ld1, c1, a2, ld2, c2, w, w_ld, w_out_major, bias, o1, old1, oc1, o2, old2, n, colstats,    \
                                num_partials, st);
  B200_NN(16, 1, 16) B200_NN(16, 1, 32) B200_NN(16, 1, 64)
  B200_NN(32, 1, 16) B200_NN(32, 1, 32) B200_NN(32, 1, 64)
  B200_NN(32, 2, 16) B200_NN(32, 2, 32) B200_NN(32, 2, 64)
#undef B200_NN
  set_error("tma_rows: unsupported shape");
  return B200_E_UNSUPPORTED;
}

// weight gradient: cout in {16, 32, 64, 128}; rows = one segment of 16 / 32 / 64 / 128 columns, or 32 + 32, or 128 + 32 (the
// decoder's concatenated inputs).  32 x 128, 64 x 128 and 128(+32) x 32 use 64-row boxes.
static int tn_plan(int c1, int c2, int cout) {  // 0 = unsupported
  if (c2 == 0) {
    if ((c1 == 16 || c1 == 32 || c1 == 64) && (cout == 16 || cout == 32 || cout == 64)) return 1;
    if ((c1 == 32 || c1 == 64) && cout == 128) return 2;
    if (c1 == 128 && cout == 32) return 3;
    return 0;
  }
  if (c1 == 32 && c2 == 32 && (cout == 16 || cout == 32 || cout == 64)) return 1;
  if (c1 == 128 && c2 == 32 && cout == 32) return 4;
  return 0;
}

bool tma_rows_tn_ok(int64_t n, const float* gy, int cout, const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2,
                    int c2) {
  if (n < 8192 || n > (int64_t)1 << 30 || !gy || !al16(gy) || !encode_fn()) return false;
  if (!seg_ok(a1, ld1, c1) || (c2 > 0 && !seg_ok(a2, ld2, c2))) return false;
  return tn_plan(c1, c2, cout) != 0;
}

int launch_tma_rows_tn(const float* gy, int cout, const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2,
                       float* gw, float* gb, int64_t n, cudaStream_t st) {
  const int ktot = c1 + c2;
  switch (tn_plan(c1, c2, cout)) {
    case 1: {
      const int ws = (c1 == 16) ? 16 : 32, ns0 = c1 / ws, ns1 = c2 / ws;
#define B200_TN(WS_, N0_, N1_, MP_) \
  if (ws == WS_ && ns0 == N0_ && ns1 == N1_ && cout == MP_) return tn_go<WS_, N0_, N1_, MP_, TR_ROWS>(gy, a1, ld1, a2, ld2, gw, ktot, gb, n, st);
      B200_TN(16, 1, 0, 16) B200_TN(16, 1, 0, 32) B200_TN(16, 1, 0, 64)
      B200_TN(32, 1, 0, 16) B200_TN(32, 1, 0, 32) B200_TN(32, 1, 0, 64)
      B200_TN(32, 2, 0, 16) B200_TN(32, 2, 0, 32) B200_TN(32, 2, 0, 64)
      B200_TN(32, 1, 1, 16) B200_TN(32, 1, 1, 32) B200_TN(32, 1, 1, 64)
#undef B200_TN
      break;
    }
    case 2:
      if (c1 == 32) return tn_go<32, 1, 0, 128, 64>(gy, a1, ld1, nullptr, 0, gw, ktot, gb, n, st);
      return tn_go<32, 2, 0, 128, 64>(gy, a1, ld1, nullptr, 0, gw, ktot, gb, n, st);
    case 3:
      return tn_go<32, 4, 0, 32, 64>(gy, a1, ld1, nullptr, 0, gw, ktot, gb, n, st);
    case 4: {  // [a1 (128) | a2 (32)]: two launches over disjoint column ranges of gw, the bias gradient rides with the first
      const int rc = tn_go<32, 4, 0, 32, 64>(gy, a1, ld1, nullptr, 0, gw, ktot, gb, n, st);
      if (rc != B200_OK) return rc;
      return tn_go<32, 1, 0, 32, TR_ROWS>(gy, a2, ld2, nullptr, 0, gw + c1, ktot, nullptr, n, st);
    }
    default:
      break;
  }
  set_error("tma_rows: unsupported shape");
  return B200_E_UNSUPPORTED;
}

}  // namespace b200
