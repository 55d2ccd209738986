// pp_util.cu -- error string, version.
#include <stdarg.h>
#include <string.h>

#include "pp_internal.cuh"

static thread_local char g_err[1024] = "";

void pp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pp_last_error(void) { return g_err; }
extern "C" const char* pp_version(void) { return "0.1.0"; }
extern "C" const char* pp_build_arch(void) { return "sm_100a"; }

// The library allocates its scratch with cudaMallocAsync.  The default pool returns freed memory
// to the driver at every synchronisation (release threshold 0), which turns the scratch arrays
// of each rebuild into fresh cudaMalloc calls; keep freed blocks cached instead.
// 256 bytes of pinned host memory per thread for the small device->host reads of the hot path (a copy
// into pageable memory is staged by the driver and costs tens of microseconds more)
void* pp_pinned_scratch() {
  static thread_local void* p = nullptr;
  if (!p && cudaHostAlloc(&p, 256, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; cudaGetLastError(); }
  return p;
}

void pp_runtime_init() {
  static thread_local int done_for = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done_for = dev;
}
