// Row-streaming Linear for the narrow layers of levels 0-1 (<= 64 input and <= 64 output channels, >= 8192 rows):
//     y[i][c] = sum_k x[i][k] * W[c][k] + b[c]        forward of torch.nn.Linear inside SharedMLP / fc0 / fc_classif
//     ga[i][k] = sum_m gy[i][m] * W[m][k]             its input gradient (the same kernel with W read the other way)
// (myria3d/models/modules/pyg_randla_net.py:42,53,97-109; the two-segment input is FPModule's torch.cat, :251.)
//
// Those layers move 13-80 MB for 0.03-0.4 GFLOP: a tile GEMM spends its time on staging, barriers and its epilogue
// (ncu, round 2: 85 warp instructions per row against a floor of 32 FFMA at 32 x 32).  Here ONE THREAD OWNS ONE ROW
// (used while padded K x padded cout <= 512, see linear_rows_ok):
//   * its K inputs sit in registers (float4 loads; a warp reads 32 consecutive rows = one contiguous run),
//   * the weights sit in shared memory as Ws[k][c]: for a fixed k every lane reads the same 16 bytes (LDS.128 broadcast,
//     one wavefront) and gets 4 FFMA out of it,
//   * the row is written back with float4 stores, again one contiguous run per warp.
// No __syncthreads in the row loop.  BatchNorm statistics (training): the warp's 32 x CO block of outputs goes through
// a private shared-memory slab and LANE c sums COLUMN c in fp64 (sum and sum of squares, like every other producer of
// b200_bn_finalize's partials: exact products, no cancellation in E[y^2] - E[y]^2); one partial row per CTA.
#include "common.cuh"

namespace b200 {

constexpr int LR_THREADS = 128;

struct LrIn {  // [a1 | a2] rows
  const float* a1;
  int64_t ld1;
  int c1;
  const float* a2;
  int64_t ld2;
  int c2;
  bool vec;  // every segment float4-addressable (16-byte aligned base, ld % 4 == 0, width % 4 == 0)
};
struct LrOut {  // columns [0, c1) -> o1, [c1, c1 + c2) -> o2; either pointer may be null (gradient not wanted)
  float* o1;
  int64_t ld1;
  int c1;
  float* o2;
  int64_t ld2;
  int c2;
  bool vec;
};

template <int KP, int CO>
__global__ void __launch_bounds__(LR_THREADS)
linear_rows_kernel(LrIn X, const float* __restrict__ w, int w_ld, bool w_out_major, const float* __restrict__ bias, LrOut Y,
                   int ktot, int cout, int64_t n, double* __restrict__ colstats) {
  __shared__ __align__(16) float Ws[KP][CO];
  __shared__ float Bs[CO];
  extern __shared__ __align__(16) float lr_dyn[];  // statistics: [warps][32][CO + 1] floats, then [warps][2 * CO] doubles
  constexpr int NW = LR_THREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto load_row = [&](int64_t row, float (&x)[KP]) {  // issue the loads of one input row (zeros beyond n / ktot)
    const bool valid = row < n;
    if (X.vec) {
#pragma unroll
      for (int k = 0; k < KP; k += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && k < ktot) {
          v = (k < X.c1) ? __ldg(reinterpret_cast<const float4*>(X.a1 + row * X.ld1 + k))
                         : __ldg(reinterpret_cast<const float4*>(X.a2 + row * X.ld2 + (k - X.c1)));
        }
        x[k] = v.x, x[k + 1] = v.y, x[k + 2] = v.z, x[k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        float v = 0.f;
        if (valid && k < ktot) v = (k < X.c1) ? __ldg(X.a1 + row * X.ld1 + k) : __ldg(X.a2 + row * X.ld2 + (k - X.c1));
        x[k] = v;
      }
    }
  };
  // the first row's loads travel while the weights are staged
  constexpr bool PREFETCH = (KP <= 32);  // register double buffering of the input row (KP = 64 would need 2 x 64 registers)
  const int64_t stride = (int64_t)gridDim.x * LR_THREADS;
  float x[KP], xn[PREFETCH ? KP : 1];
  (void)xn;
  load_row((int64_t)blockIdx.x * LR_THREADS + tid, x);

  // weights as Ws[k][c] (zero-padded): forward W is [cout][ktot] (w_out_major), the input gradient reads W[m][k'] as is
  for (int t = tid; t < KP * CO; t += LR_THREADS) {
    const int k = t / CO, c = t - k * CO;
    float v = 0.f;
    if (k < ktot && c < cout) v = w_out_major ? __ldg(w + (int64_t)c * w_ld + k) : __ldg(w + (int64_t)k * w_ld + c);
    Ws[k][c] = v;
  }
  for (int c = tid; c < CO; c += LR_THREADS) Bs[c] = (bias && c < cout) ? __ldg(bias + c) : 0.f;
  __syncthreads();

  constexpr int CL = (CO + 31) / 32;  // columns per lane in the statistics pass
  double s1[CL], s2[CL];
#pragma unroll
  for (int u = 0; u < CL; ++u) s1[u] = 0.0, s2[u] = 0.0;
  float* slab = lr_dyn + (size_t)warp * 32 * (CO + 1);

  for (int64_t base = (int64_t)blockIdx.x * LR_THREADS; base < n; base += stride) {  // CTA-uniform trip count
    const int64_t row = base + tid;
    const bool valid = row < n;
    if constexpr (PREFETCH) load_row(row + stride, xn);  // next batch of this CTA: in flight during the multiply below
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = Bs[c];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const float xk = x[k];
#pragma unroll
      for (int c = 0; c < CO; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][c]);
        ffma2_bc(xk, w4.x, w4.y, acc[c], acc[c + 1]);  // FFMA2: the row's input in the broadcast slot
        ffma2_bc(xk, w4.z, w4.w, acc[c + 2], acc[c + 3]);
      }
    }
    if (valid) {
      if (Y.vec) {
#pragma unroll
        for (int c = 0; c < CO; c += 4) {
          if (c < cout) {
            const float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
            if (c < Y.c1) {
              if (Y.o1) *reinterpret_cast<float4*>(Y.o1 + row * Y.ld1 + c) = v;
            } else if (Y.o2) {
              *reinterpret_cast<float4*>(Y.o2 + row * Y.ld2 + (c - Y.c1)) = v;
            }
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          if (c < cout) {
            if (c < Y.c1) {
              if (Y.o1) Y.o1[row * Y.ld1 + c] = acc[c];
            } else if (Y.o2) {
              Y.o2[row * Y.ld2 + (c - Y.c1)] = acc[c];
            }
          }
        }
      }
    }
    if (colstats) {
#pragma unroll
      for (int c = 0; c < CO; ++c) slab[lane * (CO + 1) + c] = valid ? acc[c] : 0.f;
      __syncwarp();
#pragma unroll
      for (int u = 0; u < CL; ++u) {
        const int c = lane + 32 * u;
        if (c < CO) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const double v = (double)slab[r * (CO + 1) + c];
            s1[u] += v;
            s2[u] = fma(v, v, s2[u]);
          }
        }
      }
      __syncwarp();
    }
    if constexpr (PREFETCH) {
#pragma unroll
      for (int k = 0; k < KP; ++k) x[k] = xn[k];
    } else {
      load_row(row + stride, x);
    }
  }

  if (colstats) {  // one partial row per CTA: [sum(0..cout) | sum of squares(0..cout)], written, never accumulated
    double* red = reinterpret_cast<double*>(lr_dyn + (size_t)NW * 32 * (CO + 1) + ((NW * 32 * (CO + 1)) & 1));
#pragma unroll
    for (int u = 0; u < CL; ++u) {
      const int c = lane + 32 * u;
      if (c < CO) {
        red[warp * 2 * CO + c] = s1[u];
        red[warp * 2 * CO + CO + c] = s2[u];
      }
    }
    __syncthreads();
    double* part = colstats + (int64_t)blockIdx.x * 2 * cout;
    for (int c = tid; c < cout; c += LR_THREADS) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int q = 0; q < NW; ++q) a += red[q * 2 * CO + c], b += red[q * 2 * CO + CO + c];
      part[c] = a;
      part[cout + c] = b;
    }
  }
}

static inline bool lr_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline int lr_pad(int v, int lo) {
  int p = lo;
  while (p < v) p <<= 1;
  return p;
}
bool linear_rows_ok(int64_t n, int k, int co) {  // k = reduction width, co = output width of the call
  // Every FFMA takes its weight from shared memory (LDS.128 broadcast = 512 bytes of register write-back for 4 FFMA per
  // lane), so the kernel tops out at one FFMA per clock and SM: it wins where the weight matrix is tiny -- 9 -> 32
  // (25 vs 31 us on 204 800 rows), 32 -> 4 (19 vs 33), 32 -> 6, 16 -> 16, 8 -> 8 -- and loses from 32 x 32 on
  // (47 vs 37 us), where the register-tiled GEMM of pointwise.cu stays (measured).
  if (n < 8192 || k < 1 || co < 1 || k > 64 || co > 64) return false;
  return lr_pad(k, 8) * lr_pad(co, 4) <= 512;
}
int linear_rows_grid(int64_t n) {  // also the number of BatchNorm-statistics partials of a forward call
  // persistent CTAs (the weights are staged once per CTA): 4 per SM, every CTA the same number of 128-row batches +- 1
  const int64_t tiles = ceil_div(n, LR_THREADS), cap = (int64_t)num_sms() * 4;
  return (int)(tiles < cap ? tiles : cap);
}

template <int KP, int CO>
static int launch_lr(const LrIn& X, const float* w, int w_ld, bool w_out_major, const float* bias, const LrOut& Y, int ktot,
                     int cout, int64_t n, double* colstats, cudaStream_t st) {
  constexpr int NW = LR_THREADS / 32;
  size_t smem = 0;
  if (colstats) smem = sizeof(float) * ((size_t)NW * 32 * (CO + 1) + 1) + sizeof(double) * NW * 2 * CO;
  auto kern = linear_rows_kernel<KP, CO>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "linear_rows smem attribute");
  }
  kern<<<linear_rows_grid(n), LR_THREADS, smem, st>>>(X, w, w_ld, w_out_major, bias, Y, ktot, cout, n, colstats);
  B200_CHECK_LAUNCH("linear_rows_kernel");
  return B200_OK;
}

template <int KP>
static int dispatch_co(const LrIn& X, const float* w, int w_ld, bool wom, const float* bias, const LrOut& Y, int ktot, int cout,
                       int64_t n, double* cs, cudaStream_t st) {
  if (cout <= 4) return launch_lr<KP, 4>(X, w, w_ld, wom, bias, Y, ktot, cout, n, cs, st);
  if (cout <= 8) return launch_lr<KP, 8>(X, w, w_ld, wom, bias, Y, ktot, cout, n, cs, st);
  if (cout <= 16) return launch_lr<KP, 16>(X, w, w_ld, wom, bias, Y, ktot, cout, n, cs, st);
  if (cout <= 32) return launch_lr<KP, 32>(X, w, w_ld, wom, bias, Y, ktot, cout, n, cs, st);
  if constexpr (KP <= 32) return launch_lr<KP, 64>(X, w, w_ld, wom, bias, Y, ktot, cout, n, cs, st);
  return B200_E_UNSUPPORTED;
}

// [o1 | o2][i][:] = [a1 | a2][i][:] . W (+ bias); W is [cout][ktot] when w_out_major (forward), else [ktot][cout]
int launch_linear_rows(const float* a1, int64_t ld1, int c1, const float* a2, int64_t ld2, int c2, const float* w, int w_ld,
                       bool w_out_major, const float* bias, float* o1, int64_t old1, int oc1, float* o2, int64_t old2, int oc2,
                       int64_t n, double* colstats, cudaStream_t st) {
  const int ktot = c1 + c2, cout = oc1 + oc2;
  LrIn X{a1, ld1, c1, a2, ld2, c2, false};
  X.vec = lr_al16(a1) && ld1 % 4 == 0 && c1 % 4 == 0 && (c2 == 0 || (lr_al16(a2) && ld2 % 4 == 0 && c2 % 4 == 0));
  LrOut Y{o1, old1, oc1, o2, old2, oc2, false};
  Y.vec = (!o1 || (lr_al16(o1) && old1 % 4 == 0)) && oc1 % 4 == 0 && (oc2 == 0 || ((!o2 || (lr_al16(o2) && old2 % 4 == 0)) && oc2 % 4 == 0));
  if (ktot <= 8) return dispatch_co<8>(X, w, w_ld, w_out_major, bias, Y, ktot, cout, n, colstats, st);
  if (ktot <= 16) return dispatch_co<16>(X, w, w_ld, w_out_major, bias, Y, ktot, cout, n, colstats, st);
  if (ktot <= 32) return dispatch_co<32>(X, w, w_ld, w_out_major, bias, Y, ktot, cout, n, colstats, st);
  return dispatch_co<64>(X, w, w_ld, w_out_major, bias, Y, ktot, cout, n, colstats, st);
}

}  // namespace b200
