// lib.cu — library-level entry points of libtae_b200.so: version, error string, device capability cache.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace tae {

static thread_local char tl_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

// The only global state: an immutable per-process capability cache (filled once).
static int g_sms = 0;
static int g_cc = 0;
static std::once_flag g_dev_once;

static void query_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    g_sms = -1;
    return;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    g_sms = -1;
    return;
  }
  g_sms = prop.multiProcessorCount;
  g_cc = prop.major * 10 + prop.minor;
}

int num_sms() {
  std::call_once(g_dev_once, query_device);
  if (g_sms <= 0) set_error("no CUDA device available");
  return g_sms;
}

// Why dynamic work lists: a persistent kernel with a static list doubles its run time as soon as one of its CTAs cannot
// become resident (NCCL's all-reduce CTAs hold a few SMs during the DDP backward): that CTA starts when the others have
// finished and then works through its whole share alone.
// -1: not decided yet (TAE_GEMM_DYNAMIC=1 / TAE_GEMM_STATIC=1 decide at first use, default static); 0 static; 1 dynamic
static std::atomic<int> g_dynamic{-1};

int set_dynamic_scheduling(int enable) { return g_dynamic.exchange(enable ? 1 : 0); }

int* sched_counter_slot() {
  constexpr int kSlots = 256;
  static int* base = nullptr;
  static std::once_flag once;
  static std::atomic<unsigned> seq{0};
  int dyn = g_dynamic.load(std::memory_order_relaxed);
  if (dyn < 0) {
    const char* e = getenv("TAE_GEMM_DYNAMIC");
    dyn = (e != nullptr && e[0] == '1') ? 1 : 0;
    int expected = -1;
    g_dynamic.compare_exchange_strong(expected, dyn);
    dyn = g_dynamic.load(std::memory_order_relaxed);
  }
  if (dyn != 1) return nullptr;
  std::call_once(once, []() {
    // one-time setup: the synchronize orders the memset (null stream) before the first launch on ANY stream
    if (cudaMalloc(&base, kSlots * 2 * sizeof(int)) != cudaSuccess || cudaMemset(base, 0, kSlots * 2 * sizeof(int)) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
      base = nullptr;
      (void)cudaGetLastError();
    }
  });
  if (base == nullptr) return nullptr;
  return base + 2 * (seq.fetch_add(1, std::memory_order_relaxed) % kSlots);
}

}  // namespace tae

extern "C" int tae_version(void) { return 1; }

// Hash of the sources this library was compiled from (tae_b200/build.py passes it): the loader refuses a library whose
// sources differ from the checkout's, because a changed `tae_gemm_args` layout would silently corrupt arguments.
#ifndef TAE_SRC_FINGERPRINT
#define TAE_SRC_FINGERPRINT "unknown"
#endif
extern "C" const char* tae_build_fingerprint(void) { return TAE_SRC_FINGERPRINT; }

extern "C" const char* tae_last_error_string(void) { return tae::tl_error; }

extern "C" int tae_num_sms(void) {
  const int n = tae::num_sms();
  return n > 0 ? n : TAE_ERR_CUDA;
}

extern "C" int tae_device_check(void) {
  if (tae::num_sms() <= 0) return TAE_ERR_CUDA;
  if (tae::g_cc != 100) {
    tae::set_error("libtae_b200 is built for sm_100a only; current device has compute capability %d.%d",
                   tae::g_cc / 10, tae::g_cc % 10);
    return TAE_ERR_UNSUPPORTED;
  }
  return TAE_OK;
}

extern "C" int tae_set_dynamic_scheduling(int enable) { return tae::set_dynamic_scheduling(enable); }

extern "C" uint64_t tae_launch_count(void) { return tae::g_launch_count.load(std::memory_order_relaxed); }
