// Error convention, launch counter and device check of the C ABI (include/b200randla.h).
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  cudaGetLastError();  // clear the sticky-less error so later calls are not poisoned
  return B200_E_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// process-wide options (b200_set_option): the only mutable global state besides the launch counter
static std::atomic<int> g_tensor_cores{1};
static std::atomic<int> g_tc_paths{0x1f};  // bit 0 linear forward, 1 linear input gradient, 2 wide weight gradient, 3 narrow weight gradient, 4 LFA
static std::atomic<int> g_tma_rows{7};  // bit 0 linear forward, 1 linear input gradient, 2 weight gradient (tma_rows.cu)
static std::atomic<int> g_bn_fused{0};   // one-launch BatchNorm backward for the small levels (pointwise.cu): measured slower, off
static std::atomic<long long*> g_tc_timeline{nullptr};
bool bn_backward_fused_enabled() { return g_bn_fused.load(std::memory_order_relaxed) != 0; }
bool tma_rows_enabled(int bit) { return (g_tma_rows.load(std::memory_order_relaxed) & bit) != 0; }
bool tensor_cores_enabled() { return g_tensor_cores.load(std::memory_order_relaxed) != 0; }
bool tc_path_enabled(int bit) {
  return tensor_cores_enabled() && (g_tc_paths.load(std::memory_order_relaxed) & bit) != 0;
}
long long* tc_debug_buffer() { return g_tc_timeline.load(std::memory_order_relaxed); }

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached_sms = v;
    cached_dev = dev;
  }
  return cached_sms;
}

}  // namespace b200

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }

const char* b200_last_error(void) { return b200::g_err; }

int64_t b200_launch_count(void) { return b200::g_launches.load(std::memory_order_relaxed); }

int b200_set_option(const char* key, int64_t value) {
  B200_REQUIRE(key, B200_E_INVALID, "b200_set_option: null key");
  if (strcmp(key, "tensor_cores") == 0) {
    b200::g_tensor_cores.store(value != 0, std::memory_order_relaxed);
    return B200_OK;
  }
  if (strcmp(key, "knn_points_per_cell") == 0) {
    b200::set_grid_points_per_cell((int)value);
    return B200_OK;
  }
  if (strcmp(key, "tensor_core_paths") == 0) {
    b200::g_tc_paths.store((int)value, std::memory_order_relaxed);
    return B200_OK;
  }
  if (strcmp(key, "tma_rows") == 0) {
    b200::g_tma_rows.store((int)value, std::memory_order_relaxed);
    return B200_OK;
  }
  if (strcmp(key, "bn_backward_fused") == 0) {
    b200::g_bn_fused.store(value != 0, std::memory_order_relaxed);
    return B200_OK;
  }
  if (strcmp(key, "tc_timeline") == 0) {
    b200::g_tc_timeline.store(reinterpret_cast<long long*>(static_cast<intptr_t>(value)), std::memory_order_relaxed);
    return B200_OK;
  }
  b200::set_error("b200_set_option: unknown option '%s' (known: tensor_cores, tensor_core_paths, tma_rows, bn_backward_fused, knn_points_per_cell, tc_timeline)", key);
  return B200_E_INVALID;
}

int64_t b200_get_option(const char* key) {
  if (key && strcmp(key, "tensor_cores") == 0) return b200::g_tensor_cores.load(std::memory_order_relaxed);
  if (key && strcmp(key, "knn_points_per_cell") == 0) return b200::get_grid_points_per_cell();
  if (key && strcmp(key, "tensor_core_paths") == 0) return b200::g_tc_paths.load(std::memory_order_relaxed);
  if (key && strcmp(key, "tma_rows") == 0) return b200::g_tma_rows.load(std::memory_order_relaxed);
  if (key && strcmp(key, "bn_backward_fused") == 0) return b200::g_bn_fused.load(std::memory_order_relaxed);
  if (key && strcmp(key, "tc_timeline") == 0)
    return static_cast<int64_t>(reinterpret_cast<intptr_t>(b200::g_tc_timeline.load(std::memory_order_relaxed)));
  return -1;
}

int b200_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b200::cuda_fail(e, "cudaGetDevice");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  B200_REQUIRE(major == 10, B200_E_UNSUPPORTED, "libb200randla needs an sm_100 device, found sm_%d%d", major, minor);
  return B200_OK;
}

}  // extern "C"
