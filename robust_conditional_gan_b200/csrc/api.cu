// Error reporting and device probing for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include <stdlib.h>

static thread_local char g_err[512] = "";

void rcgan_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#include <atomic>
static std::atomic<long> g_launches{0};
void rcgan_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool rcgan_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RCGAN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
extern "C" long rcgan_launch_count(void) { return g_launches.load(); }

extern "C" const char* rcgan_last_error(void) { return g_err; }
extern "C" int rcgan_abi_version(void) { return RCGAN_ABI_VERSION; }

// name of the conv kernel variant the last conv entry point launched on this thread (tests assert which instantiation
// a shape dispatches to: the persistent tcgen05 variants only engage above tile-count thresholds)
static thread_local char g_variant[96] = "";
// ... and every distinct variant since the last reset, ';'-separated (whole training steps: "did a persistent kernel run?")
#include <mutex>
#include <set>
#include <string>
static std::mutex g_vlog_mu;
static std::set<std::string> g_vlog;
static thread_local std::string g_vlog_out;
void rcgan_set_conv_variant(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_variant, sizeof(g_variant), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_vlog_mu);
  if (g_vlog.size() < 256) g_vlog.insert(g_variant);
}
extern "C" const char* rcgan_last_conv_variant(void) { return g_variant; }
extern "C" const char* rcgan_conv_variant_log(int reset) {
  std::lock_guard<std::mutex> lk(g_vlog_mu);
  g_vlog_out.clear();
  for (const auto& v : g_vlog) { g_vlog_out += v; g_vlog_out += ';'; }
  if (reset) g_vlog.clear();
  return g_vlog_out.c_str();
}

extern "C" int rcgan_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { rcgan_set_error("no CUDA device"); return 0; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { rcgan_set_error("cudaGetDeviceProperties failed"); return 0; }
  if (prop.major != 10) { rcgan_set_error("device is sm_%d%d; this library is sm_100a only", prop.major, prop.minor); return 0; }
  return 1;
}
