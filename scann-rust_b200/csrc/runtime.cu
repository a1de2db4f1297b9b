// Host runtime pieces shared by every entry point: thread-local error text, CUDA error mapping,
// device checks, the grow-only workspace arena.
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace scann {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

scann_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  switch (e) {
    case cudaErrorMemoryAllocation: return SCANN_RESOURCE_EXHAUSTED;
    case cudaErrorNoDevice:
    case cudaErrorInsufficientDriver:
    case cudaErrorInvalidDevice: return SCANN_UNAVAILABLE;
    default: return SCANN_INTERNAL;
  }
}

scann_status check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libscann_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return SCANN_UNAVAILABLE;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return SCANN_INVALID_ARGUMENT;
  }
  return SCANN_OK;
}

int sm_count(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
  return v;
}

int scan_acc_mode(int S) {
  // read each call (cheap) so tests can flip SCANN_ACC_MODE between searches
  const char* e = getenv("SCANN_ACC_MODE");
  int m = e ? atoi(e) : 3;
  if (m < 0 || m > 3) m = 3;
  if (m == 3 && S > 128) m = 2;  // IDP.2A packing needs per-point sums < 2^15 (128 * 255 = 32640)
  return m;
}

scann_status Workspace::reserve(size_t bytes) {
  used = 0;
  if (bytes <= cap) return SCANN_OK;
  if (base) cudaFree(base);
  base = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 8 + (1u << 20);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(reinterpret_cast<void**>(&base), want);
  }
  if (e != cudaSuccess) {
    base = nullptr;
    return cuda_fail(e, "workspace cudaMalloc", __FILE__, __LINE__);
  }
  cap = want;
  return SCANN_OK;
}

void Workspace::release() {
  if (base) cudaFree(base);
  base = nullptr;
  cap = used = 0;
}

}  // namespace scann

extern "C" {

const char* scann_last_error(void) { return scann::g_err; }

int scann_version(void) { return 100; }

scann_status scann_device_count(int* count) {
  if (!count) return SCANN_INVALID_ARGUMENT;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    scann::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return SCANN_UNAVAILABLE;
  }
  *count = n;
  return SCANN_OK;
}

}  // extern "C"
