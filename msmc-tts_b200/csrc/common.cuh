// Shared device/host helpers for the MSMC-VQ-GAN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/msmc_b200.h"

#define MSMC_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return MSMC_ERR_LAUNCH;           \
  } while (0)

#define MSMC_REQUIRE(cond)                \
  do {                                    \
    if (!(cond)) return MSMC_ERR_BAD_ARG; \
  } while (0)

namespace msmc {

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// element-wise operand transforms (msmc_xform).  `xf` is uniform per launch.  Written without a `switch`: the jump
// table (BRX, one indirect branch per element) showed up as ~1M indirect branches per launch in the weight-gradient
// producers (profiles/r01_ncu_full_wgrad128_summary.txt).  Every piece-wise linear transform is one select:
//   result = (sel > 0 ? v : s * v),  sel = v (LRELU, RELU, NONE) or aux (MUL_DLRELU, MUL_DRELU),
//   s = slope (leaky), 0 (relu) or 1 (identity); the compiler hoists sel / s selection out of the element loops.
__device__ __forceinline__ float apply_xf(int xf, float slope, float v, float aux) {
  if (xf == MSMC_XF_TANH) return tanhf(v);
  if (xf == MSMC_XF_MUL_DTANH) return v * (1.f - aux * aux);
  const float sel = (xf >= MSMC_XF_MUL_DLRELU) ? aux : v;
  const float s = (xf == MSMC_XF_NONE) ? 1.f : ((xf == MSMC_XF_LRELU || xf == MSMC_XF_MUL_DLRELU) ? slope : 0.f);
  return sel > 0.f ? v : s * v;
}
__host__ __device__ inline bool xf_needs_aux(int xf) { return xf >= MSMC_XF_MUL_DLRELU; }

// Counter-based uniform generator: one 64-bit mix (splitmix64 finaliser) per element.
// keep(mask) decisions are a pure function of (seed, salt, index) so forward and backward agree
// and a CUDA-graph replay that bumps *seed on the device draws a fresh mask.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t salt, uint64_t index) {
  uint64_t h = mix64(seed + 0x9e3779b97f4a7c15ULL * (salt + 1)) ^ (index * 0xd6e8feb86659fd93ULL);
  h = mix64(h);
  return (float)(h >> 40) * (1.0f / 16777216.0f);  // 24-bit mantissa, [0,1)
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

int num_sms();
// second pass of every weight-gradient kernel: sum the per-slice partials, scatter through the weight strides
int launch_wgrad_reduce(const msmc_conv_geom& g, const float* workspace, int splits, float* dw, float* dbias,
                        void* stream);

// persistent, warp-specialised tap-reuse convolution (conv_persist.cu); MSMC_ERR_UNSUPPORTED = shape does not fit
int conv_reuse_persistent(const msmc_conv_geom& g, const float* src, const float* src_aux, const float* wimg,
                          const float* bias, const float* residual, const float* dst_aux, float* dst,
                          int tap_stride, int n_taps, int pad_rows, int split, int BN, void* stream);

}  // namespace msmc
