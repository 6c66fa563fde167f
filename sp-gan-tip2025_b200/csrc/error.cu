#include "common.cuh"

static thread_local char g_err[512] = "";

void spgan_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* spgan_last_error(void) { return g_err; }
extern "C" int spgan_abi_version(void) { return SPGAN_ABI_VERSION; }

extern "C" int spgan_device_ok(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    spgan_set_error("no CUDA device is current");
    return 0;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    spgan_set_error("cudaGetDeviceProperties failed");
    return 0;
  }
  if (prop.major != 10) {
    spgan_set_error("libspgan_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return 0;
  }
  return 1;
}
