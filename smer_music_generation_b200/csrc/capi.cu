// Library-level plumbing of the C ABI: version, per-thread error text, device queries.
#include "common.cuh"
#include "../../include/smer_b200.h"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void smer_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// SMs the persistent kernels (GEMM, attention forward: one CTA per SM, a static work list each) may fill.  A data-parallel
// job leaves a few to the collective kernels: an all-reduce CTA cannot share an SM with a 225 KB GEMM CTA, so on a full grid
// the displaced GEMM CTAs would start only when others have finished their whole list -- the kernel takes twice as long.
static int g_reserved_sms = [] { const char* e = getenv("SMER_RESERVED_SMS"); return e ? atoi(e) : 0; }();
extern "C" int smer_set_reserved_sms(int n) {
  if (n < 0 || n > 64 || (n & 1)) { smer_set_error("smer_set_reserved_sms: need an even count in [0, 64] (got %d)", n); return SMER_ERR_ARG; }
  g_reserved_sms = n;
  return SMER_OK;
}
int smer_num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached - g_reserved_sms;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached - g_reserved_sms > 2 ? cached - g_reserved_sms : 2;
}

static int g_pdl = 0;
int smer_pdl_flag() { return g_pdl; }
extern "C" int smer_set_pdl(int on) { g_pdl = on ? 1 : 0; return SMER_OK; }

static const unsigned long long* g_seed_dev = nullptr;
const unsigned long long* smer_seed_dev() { return g_seed_dev; }
extern "C" int smer_set_seed_device_ptr(const uint64_t* p) {
  g_seed_dev = reinterpret_cast<const unsigned long long*>(p);
  return SMER_OK;
}

extern "C" int smer_version(void) { return SMER_B200_VERSION; }
extern "C" const char* smer_last_error(void) { return g_err; }

extern "C" int smer_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return major == 10 ? 1 : 0;
}
