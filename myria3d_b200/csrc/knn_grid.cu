// Exact per-cloud k-NN on a uniform 2-D bucket grid, for sm_100a.
//
// Same contract and bit-identical results as b200_knn (knn.cu) -- fp32 canonical distances, ties to
// the lower index -- but instead of scanning every point of the cloud, each query only visits the
// grid cells that can still hold one of its k nearest neighbours.  Lidar tiles are 2.5-D (50 m x 50 m
// footprint, metres of height), so the grid spans the two widest axes of each cloud and the third
// axis is handled by the distance test itself; any point distribution stays exact, only the pruning
// efficiency changes.
//
//   1. grid_meta   : one CTA per cloud -> bounding box, the two widest axes, G = f(n) cells per axis,
//                    zeroed per-cell counters
//   2. grid_count  : cell of every point (stored) + per-cell histogram (atomics)
//   3. grid_scan   : one CTA per cloud -> exclusive scan of the G*G counters (cell start offsets)
//   4. grid_scatter: counting-sort scatter into cell order as float4 (x, y, z, original index bits)
//   5. knn_grid    : square rings of cells around the query's cell are visited until the k-th distance is smaller
//                    than the distance to the border of the visited block (conservative, with slack for the
//                    float rounding of the cell assignment); a ring row is one contiguous run of the sorted
//                    array.  One thread per query; for self-queries (kNN graph) the threads follow the cell order so that
//                    a warp's lanes read (almost) the same runs.  (A warp-cooperative walk of the union block was
//                    measured: no gain, the lane-private top-k insertion dominates either way.)
//
// Insertion is lexicographic on (distance, index): the result does not depend on the visiting order,
// hence not on the (atomic, non-deterministic) order of points inside a cell.
#include <math_constants.h>

#include <atomic>

#include "common.cuh"

namespace b200 {

constexpr int GRID_THREADS = 128;
constexpr int GRID_MAX_G = 128;
constexpr float GRID_POINTS_PER_CELL = 6.0f;

struct GridMeta {
  float oa, ob;      // grid origin along the two grid axes
  float inva, invb;  // cells per unit length
  float wa, wb;      // cell width
  float slack;       // absolute rounding slack of the block-border distance
  int da, db;        // which coordinates (0..2) span the grid
  int g;             // cells per axis
  int pad0, pad1;
};

// average points per (2-D) cell; b200_set_option("knn_points_per_cell", v) overrides the default for tuning runs
static std::atomic<int> g_grid_ppc{0};  // 0 = automatic (by k, below)
void set_grid_points_per_cell(int v) { g_grid_ppc.store(v < 0 ? 0 : (v > 64 ? 64 : v), std::memory_order_relaxed); }
int get_grid_points_per_cell() { return g_grid_ppc.load(std::memory_order_relaxed); }
// Measured (scripts/knn_ppc_sweep.py, 16 x 12 800 points k = 16 / 4 x 65 536 points k = 32): 6 points per cell 496 / 1605 us,
// 8: 468 / 1523, 16: 458 / 1407 -- fewer, fuller cells mean fewer half-empty candidate batches per ring row.
static float grid_ppc_for(int k) {
  const int v = get_grid_points_per_cell();
  if (v > 0) return (float)v;
  return k >= 24 ? 16.f : (k >= 8 ? 8.f : GRID_POINTS_PER_CELL);
}

__host__ __device__ __forceinline__ int grid_cells_for(long long n, float ppc) {
  int g = (int)sqrtf((float)n / ppc);
  if (g < 1) g = 1;
  if (g > GRID_MAX_G) g = GRID_MAX_G;
  return g;
}

__device__ __forceinline__ int grid_cell_1d(float p, float o, float inv, int g) {
  int c = (int)((p - o) * inv);
  c = c < 0 ? 0 : c;
  return c > g - 1 ? g - 1 : c;
}

// ---- 1. per-cloud bounding box and grid geometry
__global__ void __launch_bounds__(256)
grid_meta_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr, GridMeta* __restrict__ meta,
                 int* __restrict__ counts, int stride, float ppc) {
  const int cloud = blockIdx.x;
  const int64_t xs = ptr[cloud], xe = ptr[cloud + 1];
  float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
  float hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int64_t i = xs + threadIdx.x; i < xe; i += blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = pos[3 * i + d];
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  }
  __shared__ float slo[8][3], shi[8][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) slo[warp][d] = lo[d], shi[warp][d] = hi[d];
  }
  __syncthreads();
  const int g = grid_cells_for(xe - xs, ppc);
  if (threadIdx.x == 0) {
    float ext[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      for (int w = 1; w < 8; ++w) {
        slo[0][d] = fminf(slo[0][d], slo[w][d]);
        shi[0][d] = fmaxf(shi[0][d], shi[w][d]);
      }
      ext[d] = (xe > xs) ? (shi[0][d] - slo[0][d]) : 0.f;
    }
    // the two widest axes (ties: lower axis first)
    int da = 0, db = 1;
    {
      int small = 0;
      if (ext[1] < ext[small]) small = 1;
      if (ext[2] < ext[small]) small = 2;
      da = (small == 0) ? 1 : 0;
      db = (small == 2) ? 1 : 2;
    }
    GridMeta m;
    m.da = da, m.db = db, m.g = g, m.pad0 = 0, m.pad1 = 0;
    m.oa = (xe > xs) ? slo[0][da] : 0.f;
    m.ob = (xe > xs) ? slo[0][db] : 0.f;
    m.inva = (ext[da] > 0.f) ? (float)g / ext[da] : 0.f;
    m.invb = (ext[db] > 0.f) ? (float)g / ext[db] : 0.f;
    m.wa = (ext[da] > 0.f) ? ext[da] / (float)g : CUDART_INF_F;
    m.wb = (ext[db] > 0.f) ? ext[db] / (float)g : CUDART_INF_F;
    const float scale = fmaxf(fmaxf(fabsf(slo[0][da]), fabsf(shi[0][da])), fmaxf(fabsf(slo[0][db]), fabsf(shi[0][db])));
    m.slack = 1e-5f * (scale + ext[da] + ext[db]);
    meta[cloud] = m;
  }
  for (int c = threadIdx.x; c < g * g; c += blockDim.x) counts[(int64_t)cloud * stride + c] = 0;
}

// ---- 2. histogram
__global__ void __launch_bounds__(256)
grid_count_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr, const GridMeta* __restrict__ meta,
                  int* __restrict__ counts, int* __restrict__ cell_of, int stride) {
  const int cloud = blockIdx.y;
  const int64_t xs = ptr[cloud], xe = ptr[cloud + 1];
  const int64_t i = xs + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= xe) return;
  const GridMeta m = meta[cloud];
  const float pa = pos[3 * i + m.da], pb = pos[3 * i + m.db];  // (global loads: dynamic offsets are fine)
  const int cell = grid_cell_1d(pb, m.ob, m.invb, m.g) * m.g + grid_cell_1d(pa, m.oa, m.inva, m.g);
  cell_of[i] = cell;
  atomicAdd(&counts[(int64_t)cloud * stride + cell], 1);
}

// ---- 3. exclusive scan of the cell counters (one CTA per cloud); counters become scatter cursors
__global__ void __launch_bounds__(256)
grid_scan_kernel(const int64_t* __restrict__ ptr, int* __restrict__ counts, int* __restrict__ starts, int stride, float ppc) {
  const int cloud = blockIdx.x;
  const int g = grid_cells_for(ptr[cloud + 1] - ptr[cloud], ppc);
  const int cells = g * g;
  int* cnt = counts + (int64_t)cloud * stride;
  int* st = starts + (int64_t)cloud * (stride + 1);
  const int per = (cells + 255) / 256;
  const int c0 = threadIdx.x * per;
  int local = 0;
  for (int c = c0; c < c0 + per && c < cells; ++c) local += cnt[c];
  __shared__ int part[256];
  part[threadIdx.x] = local;
  __syncthreads();
  // Hillis-Steele inclusive scan over 256 partial sums
  for (int o = 1; o < 256; o <<= 1) {
    const int v = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - local;
  for (int c = c0; c < c0 + per && c < cells; ++c) {
    const int v = cnt[c];
    st[c] = run;
    cnt[c] = run;  // cursor for the scatter pass
    run += v;
  }
  if (threadIdx.x == 255) st[cells] = part[255];
}

// ---- 4. counting-sort scatter
__global__ void __launch_bounds__(256)
grid_scatter_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr, int* __restrict__ cursors,
                    const int* __restrict__ cell_of, float4* __restrict__ sorted, int stride) {
  const int cloud = blockIdx.y;
  const int64_t xs = ptr[cloud], xe = ptr[cloud + 1];
  const int64_t i = xs + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= xe) return;
  const int p = atomicAdd(&cursors[(int64_t)cloud * stride + cell_of[i]], 1);
  sorted[xs + p] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __int_as_float((int)i));
}

// ---- 5. ring search
template <int KMAX>
struct TopKLex {
  float d[KMAX];
  int idx[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) d[j] = CUDART_INF_F, idx[j] = -1;
  }
  __device__ __forceinline__ bool after(int j, float dist, int i) const {  // slot j sorts after (dist, i)
    return d[j] > dist || (d[j] == dist && idx[j] > i);
  }
  __device__ __forceinline__ void push(float dist, int i) {
    if (after(KMAX - 1, dist, i)) {
#pragma unroll
      for (int j = KMAX - 1; j >= 1; --j) {
        if (after(j - 1, dist, i)) {
          d[j] = d[j - 1];
          idx[j] = idx[j - 1];
        } else if (after(j, dist, i)) {
          d[j] = dist;
          idx[j] = i;
        }
      }
      if (after(0, dist, i)) d[0] = dist, idx[0] = i;
    }
  }
};

template <int KMAX, bool SELF>
__global__ void __launch_bounds__(GRID_THREADS, (KMAX <= 16 ? 4 : 1))
knn_grid_kernel(const float4* __restrict__ sorted, const int* __restrict__ starts, const GridMeta* __restrict__ meta,
                const int64_t* __restrict__ ptr_x, const float* __restrict__ pos_y, const int64_t* __restrict__ ptr_y,
                int stride, int k, int kt, int32_t* __restrict__ nbr, float* __restrict__ dist2) {
  const int cloud = blockIdx.y;
  const int64_t xs = ptr_x[cloud], xe = ptr_x[cloud + 1];
  const int64_t ys = SELF ? xs : ptr_y[cloud], ye = SELF ? xe : ptr_y[cloud + 1];
  const int64_t t = ys + (int64_t)blockIdx.x * GRID_THREADS + threadIdx.x;
  if (t >= ye) return;

  float q[3];
  int64_t out_row;
  if (SELF) {
    const float4 s = sorted[t];  // t-th point of the cloud in cell order
    q[0] = s.x, q[1] = s.y, q[2] = s.z;
    out_row = __float_as_int(s.w);
  } else {
    q[0] = pos_y[3 * t], q[1] = pos_y[3 * t + 1], q[2] = pos_y[3 * t + 2];
    out_row = t;
  }
  TopKLex<KMAX> top;
  top.init();

  if (xe > xs) {
    const GridMeta m = meta[cloud];
    const int g = m.g;
    const int* st = starts + (int64_t)cloud * (stride + 1);
    const float4* pts = sorted + xs;
    const float qa = (m.da == 0) ? q[0] : ((m.da == 1) ? q[1] : q[2]);  // static indexing: q stays in registers
    const float qb = (m.db == 0) ? q[0] : ((m.db == 1) ? q[1] : q[2]);
    const int ca = grid_cell_1d(qa, m.oa, m.inva, g), cb = grid_cell_1d(qb, m.ob, m.invb, g);

    for (int r = 0;; ++r) {
      const int a0 = ca - r, a1 = ca + r, b0 = cb - r, b1 = cb + r;
      const int alo = a0 < 0 ? 0 : a0, ahi = a1 > g - 1 ? g - 1 : a1;
      const int blo = b0 < 0 ? 0 : b0, bhi = b1 > g - 1 ? g - 1 : b1;
      for (int bb = blo; bb <= bhi; ++bb) {
        // Every lane runs the SAME two candidate loops with its own ranges (empty where not needed): run 1 = the whole
        // ring row (a contiguous run of the sorted array) or the left cell of an interior row, run 2 = the right cell of
        // an interior row.  Separate call sites per case made the lanes of a warp -- neighbouring queries whose rings
        // are shifted by a cell -- take different code paths: 9 of 32 lanes active in the distance evaluation
        // (profiles/ncu_knn_grid_r01.summary.txt).
        const bool full = (bb == b0 || bb == b1);
        const int* row = st + bb * g;
        int s1 = 0, e1 = 0, s2 = 0, e2 = 0;
        if (full) {
          s1 = row[alo], e1 = row[ahi + 1];
        } else {
          if (a0 >= 0) s1 = row[a0], e1 = row[a0 + 1];
          if (a1 <= g - 1) s2 = row[a1], e2 = row[a1 + 1];
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          const int s = pass ? s2 : s1, e = pass ? e2 : e1;
          for (int u = s; u < e; ++u) {
            const float4 c = __ldg(pts + u);
            top.push(dist2_rn(c.x, c.y, c.z, q[0], q[1], q[2]), __float_as_int(c.w));
          }
        }
      }
      if (a0 <= 0 && a1 >= g - 1 && b0 <= 0 && b1 >= g - 1) break;  // the whole grid has been visited
      float kth = CUDART_INF_F;  // k-th best so far (static indexing keeps the list in registers)
#pragma unroll
      for (int j = 0; j < KMAX; ++j)
        if (j == k - 1) kth = top.d[j];
      if (kth < CUDART_INF_F) {
        // every unvisited point lies outside the block [a0, a1] x [b0, b1]: lower-bound its distance
        float mind = CUDART_INF_F;
        if (a0 > 0) mind = fminf(mind, qa - (m.oa + (float)a0 * m.wa));
        if (a1 < g - 1) mind = fminf(mind, (m.oa + (float)(a1 + 1) * m.wa) - qa);
        if (b0 > 0) mind = fminf(mind, qb - (m.ob + (float)b0 * m.wb));
        if (b1 < g - 1) mind = fminf(mind, (m.ob + (float)(b1 + 1) * m.wb) - qb);
        mind -= m.slack;
        if (mind > 0.f && mind * mind > kth) break;
      }
    }
  }

  int32_t* orow = nbr + out_row * kt;
  float* drow = dist2 ? dist2 + out_row * kt : nullptr;
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

// ---- 5b. warp-cooperative ring search (production path for the kNN graph: kt = 16 or 32 neighbour slots)
// A group of LANES = kt lanes (half a warp for K = 16) serves one query.  The group keeps the current top-k as a
// sorted list DISTRIBUTED over its lanes (lane l holds the l-th best (distance, index) pair); candidates are read LANES
// at a time with one coalesced load per group, filtered against the k-th best pair with a ballot, and every survivor
// is inserted with one ballot + two shuffles (rank = number of list entries sorting before it; the lanes behind shift up
// by one).  This is the "warp-shuffle top-k": no data-dependent per-thread insertion loop, every lane does useful work
// in every instruction (round 1's thread-per-query kernel ran 7.6 of 32 lanes).  All loops have warp-uniform trip
// counts (ring rows by offset, batches by the longer of the two groups' runs), so the two half-warp groups of a warp --
// in cell order they usually sit in the same cell -- never diverge and full-mask shuffles are legal throughout.
// Ordering is the same lexicographic (distance, index) rule as everywhere else: results are bit-identical to knn.cu.
template <int LANES, bool SELF>
__global__ void __launch_bounds__(GRID_THREADS)
knn_grid_warp_kernel(const float4* __restrict__ sorted, const int* __restrict__ starts, const GridMeta* __restrict__ meta,
                     const int64_t* __restrict__ ptr_x, const float* __restrict__ pos_y, const int64_t* __restrict__ ptr_y,
                     int stride, int k, int kt, int32_t* __restrict__ nbr, float* __restrict__ dist2) {
  constexpr int QPW = 32 / LANES, QPB = (GRID_THREADS / 32) * QPW;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned GMASK = (LANES == 32) ? 0xffffffffu : ((1u << (LANES & 31)) - 1u);
  const int cloud = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane % LANES, gshift = (lane / LANES) * LANES;
  const int64_t xs = ptr_x[cloud], xe = ptr_x[cloud + 1];
  const int64_t ys = SELF ? xs : ptr_y[cloud], ye = SELF ? xe : ptr_y[cloud + 1];
  const int64_t t = ys + (int64_t)blockIdx.x * QPB + warp * QPW + lane / LANES;
  const bool active = t < ye;  // group-uniform
  if (!__any_sync(FULL, active)) return;

  float q0 = 0.f, q1 = 0.f, q2 = 0.f;
  int64_t out_row = 0;
  if (active) {
    if (SELF) {
      const float4 s = sorted[t];  // t-th point of the cloud in cell order
      q0 = s.x, q1 = s.y, q2 = s.z;
      out_row = __float_as_int(s.w);
    } else {
      q0 = pos_y[3 * t], q1 = pos_y[3 * t + 1], q2 = pos_y[3 * t + 2];
      out_row = t;
    }
  }
  float d = CUDART_INF_F;  // this lane's entry of the group's sorted top list
  int id = -1;

  if (xe > xs) {
    const GridMeta m = meta[cloud];
    const int g = m.g;
    const int* st = starts + (int64_t)cloud * (stride + 1);
    const float4* pts = sorted + xs;
    const float qa = (m.da == 0) ? q0 : ((m.da == 1) ? q1 : q2);
    const float qb = (m.db == 0) ? q0 : ((m.db == 1) ? q1 : q2);
    const int ca = grid_cell_1d(qa, m.oa, m.inva, g), cb = grid_cell_1d(qb, m.ob, m.invb, g);
    bool done = !active;

    for (int r = 0;; ++r) {
      const int a0 = ca - r, a1 = ca + r, b0 = cb - r, b1 = cb + r;
      const int alo = a0 < 0 ? 0 : a0, ahi = a1 > g - 1 ? g - 1 : a1;
      for (int db = -r; db <= r; ++db) {
        const int bb = cb + db;
        const bool row_ok = !done && bb >= 0 && bb <= g - 1;
        const bool full_row = (db == -r || db == r);
        // run 1 = the whole ring row (contiguous in the sorted array) or the left cell of an interior row,
        // run 2 = the right cell of an interior row
        int s1 = 0, e1 = 0, s2 = 0, e2 = 0;
        if (row_ok) {
          const int* row = st + bb * g;
          if (full_row) {
            s1 = row[alo], e1 = row[ahi + 1];
          } else {
            if (a0 >= 0) s1 = row[a0], e1 = row[a0 + 1];
            if (a1 <= g - 1) s2 = row[a1], e2 = row[a1 + 1];
          }
        }
        {
          // the two runs of an interior row (left and right cell) are walked as ONE concatenated run: half as many
          // candidate batches there, fuller lanes
          const int len1 = e1 - s1, mine = len1 + (e2 - s2);
          int len = mine;
          if (LANES < 32) len = max(len, __shfl_xor_sync(FULL, len, 16));  // the longer of the two groups' runs
          for (int u0 = 0; u0 < len; u0 += LANES) {
            const int uu = u0 + lg;
            const int u = (uu < len1) ? (s1 + uu) : (s2 + (uu - len1));
            const bool valid = uu < mine;
            float dc = CUDART_INF_F;
            int ic = 0x7fffffff;
            if (valid) {
              const float4 c = __ldg(pts + u);
              dc = dist2_rn(c.x, c.y, c.z, q0, q1, q2);
              ic = __float_as_int(c.w);
            }
            const float kd = __shfl_sync(FULL, d, k - 1, LANES);
            const int ki = __shfl_sync(FULL, id, k - 1, LANES);
            const bool passes = valid && (dc < kd || (dc == kd && (ki < 0 || ic < ki)));
            unsigned mm = (__ballot_sync(FULL, passes) >> gshift) & GMASK;
            while (__any_sync(FULL, mm != 0)) {
              const bool has = mm != 0;
              const int src = has ? (__ffs(mm) - 1) : 0;
              mm &= mm - 1;
              const float bd = __shfl_sync(FULL, dc, src, LANES);
              const int bi = __shfl_sync(FULL, ic, src, LANES);
              // rank of the candidate = list entries sorting before it (empty slots: distance +inf, never before)
              const bool before = (d < bd) || (d == bd && id >= 0 && id < bi);
              const int pos = __popc((__ballot_sync(FULL, before) >> gshift) & GMASK);
              const float pd = __shfl_up_sync(FULL, d, 1, LANES);
              const int pi = __shfl_up_sync(FULL, id, 1, LANES);
              if (has && pos < k) {
                if (lg > pos) {
                  d = pd, id = pi;
                } else if (lg == pos) {
                  d = bd, id = bi;
                }
              }
            }
          }
        }
      }
      if (a0 <= 0 && a1 >= g - 1 && b0 <= 0 && b1 >= g - 1) done = true;  // the whole grid has been visited
      const float kth = __shfl_sync(FULL, d, k - 1, LANES);  // (outside the branch: both groups execute it together)
      if (!done && kth < CUDART_INF_F) {
        // every unvisited point lies outside the block [a0, a1] x [b0, b1]: lower-bound its distance
        float mind = CUDART_INF_F;
        if (a0 > 0) mind = fminf(mind, qa - (m.oa + (float)a0 * m.wa));
        if (a1 < g - 1) mind = fminf(mind, (m.oa + (float)(a1 + 1) * m.wa) - qa);
        if (b0 > 0) mind = fminf(mind, qb - (m.ob + (float)b0 * m.wb));
        if (b1 < g - 1) mind = fminf(mind, (m.ob + (float)(b1 + 1) * m.wb) - qb);
        mind -= m.slack;
        if (mind > 0.f && mind * mind > kth) done = true;
      }
      if (__all_sync(FULL, done)) break;
    }
  }

  if (active && lg < kt) {
    const bool keep = (lg < k) && (id >= 0);
    nbr[out_row * kt + lg] = keep ? id : -1;
    if (dist2) dist2[out_row * kt + lg] = keep ? d : CUDART_INF_F;
  }
}

struct GridWorkspace {
  GridMeta* meta;
  int* counts;
  int* starts;
  int* cell_of;
  float4* sorted;
  size_t bytes;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static GridWorkspace carve(void* base, int64_t nx, int32_t num_clouds, int stride) {
  GridWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.meta = static_cast<GridMeta*>(take(sizeof(GridMeta) * (size_t)num_clouds));
  w.counts = static_cast<int*>(take(sizeof(int) * (size_t)num_clouds * stride));
  w.starts = static_cast<int*>(take(sizeof(int) * (size_t)num_clouds * (stride + 1)));
  w.cell_of = static_cast<int*>(take(sizeof(int) * (size_t)nx));
  w.sorted = static_cast<float4*>(take(sizeof(float4) * (size_t)nx));
  w.bytes = off;
  return w;
}

template <int KMAX>
static int launch_grid_search(bool self, const GridWorkspace& w, const int64_t* ptr_x, const float* pos_y,
                              const int64_t* ptr_y, int stride, int num_clouds, int64_t max_q, int k, int kt,
                              int32_t* nbr, float* dist2, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(max_q, GRID_THREADS), (unsigned)num_clouds);
  if (self)
    knn_grid_kernel<KMAX, true><<<grid, GRID_THREADS, 0, st>>>(w.sorted, w.starts, w.meta, ptr_x, pos_y, ptr_y, stride, k, kt, nbr, dist2);
  else
    knn_grid_kernel<KMAX, false><<<grid, GRID_THREADS, 0, st>>>(w.sorted, w.starts, w.meta, ptr_x, pos_y, ptr_y, stride, k, kt, nbr, dist2);
  B200_CHECK_LAUNCH("knn_grid_kernel");
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_knn_grid_workspace_bytes(int64_t nx, int32_t num_clouds, int64_t max_x_per_cloud) {
  if (nx < 0 || num_clouds < 0 || max_x_per_cloud < 0) return -1;
  // (sized for the finest grid any k may ask for)
  const int v = get_grid_points_per_cell();
  const int g = grid_cells_for(max_x_per_cloud, v > 0 ? (float)v : GRID_POINTS_PER_CELL);
  return (int64_t)carve(nullptr, nx, num_clouds, g * g).bytes;
}

extern "C" int b200_knn_grid(const float* pos_x, const int64_t* ptr_x, int64_t nx, const float* pos_y,
                             const int64_t* ptr_y, int64_t ny, int32_t num_clouds, int64_t max_x_per_cloud,
                             int64_t max_y_per_cloud, int32_t k, int32_t kt, int32_t* nbr, float* dist2,
                             void* workspace, int64_t workspace_bytes, void* stream) {
  B200_REQUIRE(pos_x && ptr_x && pos_y && ptr_y && nbr && workspace, B200_E_INVALID, "b200_knn_grid: null pointer");
  B200_REQUIRE(k >= 1 && kt >= k, B200_E_INVALID, "b200_knn_grid: need 1 <= k <= kt (k=%d kt=%d)", k, kt);
  B200_REQUIRE(k <= 64, B200_E_UNSUPPORTED, "b200_knn_grid: k=%d > 64 not supported", k);
  B200_REQUIRE(nx < (int64_t(1) << 31) && ny < (int64_t(1) << 31), B200_E_UNSUPPORTED, "b200_knn_grid: more than 2^31 points");
  B200_REQUIRE(num_clouds >= 0 && num_clouds <= 65535, B200_E_UNSUPPORTED, "b200_knn_grid: num_clouds=%d out of [0,65535]", num_clouds);
  B200_REQUIRE(((uintptr_t)workspace & 255) == 0, B200_E_INVALID, "b200_knn_grid: workspace must be 256-byte aligned");
  if (ny == 0 || num_clouds == 0 || max_y_per_cloud <= 0) return B200_OK;
  const float ppc = grid_ppc_for(k);
  const int g = grid_cells_for(max_x_per_cloud, ppc);
  const int stride = g * g;
  const GridWorkspace w = carve(workspace, nx, num_clouds, stride);
  B200_REQUIRE((int64_t)w.bytes <= workspace_bytes, B200_E_INVALID, "b200_knn_grid: workspace too small (%lld < %lld)",
               (long long)workspace_bytes, (long long)w.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool self = (pos_x == pos_y) && (ptr_x == ptr_y) && (nx == ny);

  grid_meta_kernel<<<(unsigned)num_clouds, 256, 0, st>>>(pos_x, ptr_x, w.meta, w.counts, stride, ppc);
  B200_CHECK_LAUNCH("grid_meta_kernel");
  if (max_x_per_cloud > 0) {
    dim3 pgrid((unsigned)ceil_div(max_x_per_cloud, 256), (unsigned)num_clouds);
    grid_count_kernel<<<pgrid, 256, 0, st>>>(pos_x, ptr_x, w.meta, w.counts, w.cell_of, stride);
    B200_CHECK_LAUNCH("grid_count_kernel");
    grid_scan_kernel<<<(unsigned)num_clouds, 256, 0, st>>>(ptr_x, w.counts, w.starts, stride, ppc);
    B200_CHECK_LAUNCH("grid_scan_kernel");
    grid_scatter_kernel<<<pgrid, 256, 0, st>>>(pos_x, ptr_x, w.counts, w.cell_of, w.sorted, stride);
    B200_CHECK_LAUNCH("grid_scatter_kernel");
  } else {
    grid_scan_kernel<<<(unsigned)num_clouds, 256, 0, st>>>(ptr_x, w.counts, w.starts, stride, ppc);
    B200_CHECK_LAUNCH("grid_scan_kernel");
  }
  if ((kt == 16 || kt == 32) && k > 1) {  // kNN graph / k = 10 interpolation: warp-cooperative search
    const int qpb = (GRID_THREADS / 32) * (32 / kt);
    dim3 grid((unsigned)ceil_div(max_y_per_cloud, qpb), (unsigned)num_clouds);
#define B200_WARP_CASE(L, S) \
    knn_grid_warp_kernel<L, S><<<grid, GRID_THREADS, 0, st>>>(w.sorted, w.starts, w.meta, ptr_x, pos_y, ptr_y, stride, k, kt, nbr, dist2)
    if (kt == 16 && self) B200_WARP_CASE(16, true);
    else if (kt == 16) B200_WARP_CASE(16, false);
    else if (self) B200_WARP_CASE(32, true);
    else B200_WARP_CASE(32, false);
#undef B200_WARP_CASE
    B200_CHECK_LAUNCH("knn_grid_warp_kernel");
    return B200_OK;
  }
#define B200_KNN_CASE(KM) \
  if (k <= KM) return launch_grid_search<KM>(self, w, ptr_x, pos_y, ptr_y, stride, num_clouds, max_y_per_cloud, k, kt, nbr, dist2, st)
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
