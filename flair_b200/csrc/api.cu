// Library-wide plumbing of the C ABI: error string, device check, driver entry
// points.  See include/flair_b200.h for the conventions.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {
thread_local char g_err[1024] = "";
}

void flair_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int flair_num_sms() {
  static std::atomic<int> sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int v = sms[dev & 63].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    sms[dev & 63].store(v, std::memory_order_relaxed);
  }
  return v;
}

int flair_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FLAIR_PDL");
    // on by default since round 2 (FLAIR_PDL=0 opts out): with the weight prefetch ahead of griddepcontrol.wait the
    // serial BasicVSR++ chain gains 3-5 % inside CUDA graphs (profiles/r02_summary.md); all GPU tests run with it
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v;
}

flair_tmap_encode_fn flair_get_tmap_encode() {
  static flair_tmap_encode_fn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<flair_tmap_encode_fn>(sym);
    }
  }
  return fn;
}

extern "C" const char* flair_last_error(void) { return g_err; }
extern "C" int flair_version(void) { return FLAIR_B200_VERSION; }

extern "C" int flair_check_device(int dev) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    flair_set_error("flair_check_device: no CUDA device visible (this library has no CPU path)");
    return FLAIR_ERR_UNSUPPORTED;
  }
  if (dev < 0 || dev >= count) {
    flair_set_error("flair_check_device: device %d out of range (0..%d)", dev, count - 1);
    return FLAIR_ERR_INVALID;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    flair_set_error("flair_check_device: device %d is sm_%d%d; kernels are built for sm_100a only",
                    dev, major, minor);
    return FLAIR_ERR_UNSUPPORTED;
  }
  return 0;
}
