#include "common.cuh"
namespace msmc {
int num_sms() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;  // B200
  }
  cached = n;
  return n;
}
}  // namespace msmc
extern "C" int msmc_version(void) { return 100; }
extern "C" int msmc_num_sms(void) { return msmc::num_sms(); }
