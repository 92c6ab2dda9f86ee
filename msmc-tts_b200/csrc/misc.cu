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

// ------------------------------------------------------------------------------------------------
// Fused multi-tensor Adam / AdamW step with the global gradient-norm clip folded in (reference
// trainers/msmctts_trainer.py:203-207 clip_grad_norm_ + optimizers/__init__.py:53-78 step).  torch's capturable
// foreach path spends one tiny kernel per parameter on the device-resident step size (~1.2k launches per train
// step here) plus a dozen multi-tensor passes, and clip_grad_norm_ another ~10; this is TWO launches per optimizer:
//   adam_prepare_kernel : per 16k-element chunk the sum of squares of the gradient (fixed order) -> partial[chunk];
//                         the first chunk of every tensor also advances THAT tensor's step counter (torch advances
//                         a parameter's step only when it has a gradient: counters are per parameter).
//   adam_multi_kernel   : every block re-reduces partial[] in a fixed order (a few KB from L2) into the total norm,
//                         forms clip = min(1, max_norm / (norm + 1e-6)) exactly like torch.nn.utils.clip_grad_norm_,
//                         and applies the update with the clipped gradient (written back, as clip_grad_norm_ does).
// Pointers come from a device table; lr, the step counters and the norm are device scalars, so the launches are
// CUDA-graph replayable.
// ------------------------------------------------------------------------------------------------
namespace msmc {
namespace {
constexpr int ADAM_CHUNK = 16384;
constexpr int ADAM_THREADS = 256;

__device__ __forceinline__ float adam_block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (ADAM_THREADS / 32) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_prepare_kernel(
    const unsigned long long* __restrict__ table, int n_tensors, const long long* __restrict__ sizes,
    const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_index, const int* __restrict__ step_index,
    float* __restrict__ steps, float* __restrict__ partial, int want_norm) {
  __shared__ float sh[ADAM_THREADS / 32];
  const int t = chunk_tensor[blockIdx.x];
  const int ci = chunk_index[blockIdx.x];
  if (ci == 0 && threadIdx.x == 0) steps[step_index[t]] += 1.f;
  if (!want_norm) return;
  const long long beg = (long long)ci * ADAM_CHUNK;
  const long long n = sizes[t];
  const long long end = beg + ADAM_CHUNK < n ? beg + ADAM_CHUNK : n;
  const float* __restrict__ g = reinterpret_cast<const float*>(table[n_tensors + t]);
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long end4 = beg + ((end - beg) & ~3LL);
    for (long long i = beg + 4LL * threadIdx.x; i < end4; i += 4LL * ADAM_THREADS) {
      const float4 x = *reinterpret_cast<const float4*>(g + i);
      acc += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    }
    for (long long i = end4 + threadIdx.x; i < end; i += ADAM_THREADS) acc += g[i] * g[i];
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += ADAM_THREADS) acc += g[i] * g[i];
  }
  const float tot = adam_block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_multi_kernel(
    const unsigned long long* __restrict__ table, int n_tensors, const long long* __restrict__ sizes,
    const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_index, const int* __restrict__ step_index,
    const float* __restrict__ steps, const float* __restrict__ partial, int n_chunks, float max_norm,
    float* __restrict__ norm_out, const float* __restrict__ lr_p, float beta1, float beta2, float eps,
    float weight_decay, int decoupled) {
  __shared__ float sh[ADAM_THREADS / 32];
  __shared__ float s_clip;
  const int t = chunk_tensor[blockIdx.x];
  const long long beg = (long long)chunk_index[blockIdx.x] * ADAM_CHUNK;
  const long long n = sizes[t];
  const long long end = beg + ADAM_CHUNK < n ? beg + ADAM_CHUNK : n;
  float* __restrict__ p = reinterpret_cast<float*>(table[t]);
  float* __restrict__ g = reinterpret_cast<float*>(table[n_tensors + t]);
  float* __restrict__ m = reinterpret_cast<float*>(table[2 * n_tensors + t]);
  float* __restrict__ v = reinterpret_cast<float*>(table[3 * n_tensors + t]);
  float clip = 1.f;
  if (partial != nullptr) {
    // identical order in every block -> every block derives the identical coefficient
    float acc = 0.f;
    for (int i = threadIdx.x; i < n_chunks; i += ADAM_THREADS) acc += partial[i];
    const float tot = adam_block_sum(acc, sh);
    if (threadIdx.x == 0) {
      const float norm = sqrtf(tot);
      const float c = max_norm / (norm + 1e-6f);
      s_clip = c < 1.f ? c : 1.f;
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
    __syncthreads();
    clip = s_clip;
  }
  const bool write_g = partial != nullptr;
  const float lr = *lr_p, step = steps[step_index[t]];
  const float bc1 = 1.f - powf(beta1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * weight_decay;
  auto upd = [&](float& pw, float& gw_io, float& mw, float& vw) {
    float gw = gw_io * clip;                    // clip_grad_norm_: grad.mul_(clip_coef_clamped)
    gw_io = gw;
    if (decoupled) pw *= decay;                 // AdamW: param.mul_(1 - lr * wd)
    else gw = fmaf(weight_decay, pw, gw);       // Adam : grad + wd * param
    mw = mw + (gw - mw) * (1.f - beta1);        // exp_avg.lerp_(grad, 1 - beta1)
    vw = fmaf(vw, beta2, (1.f - beta2) * gw * gw);
    const float denom = sqrtf(vw) / bc2_sqrt + eps;
    pw -= step_size * (mw / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const long long end4 = beg + ((end - beg) & ~3LL);
    for (long long i = beg + 4LL * threadIdx.x; i < end4; i += 4LL * ADAM_THREADS) {
      float4 pw = *reinterpret_cast<float4*>(p + i);
      float4 gw = *reinterpret_cast<const float4*>(g + i);
      float4 mw = *reinterpret_cast<float4*>(m + i);
      float4 vw = *reinterpret_cast<float4*>(v + i);
      upd(pw.x, gw.x, mw.x, vw.x); upd(pw.y, gw.y, mw.y, vw.y);
      upd(pw.z, gw.z, mw.z, vw.z); upd(pw.w, gw.w, mw.w, vw.w);
      *reinterpret_cast<float4*>(p + i) = pw;
      *reinterpret_cast<float4*>(m + i) = mw;
      *reinterpret_cast<float4*>(v + i) = vw;
      if (write_g) *reinterpret_cast<float4*>(g + i) = gw;
    }
    for (long long i = end4 + threadIdx.x; i < end; i += ADAM_THREADS) {
      float gw = g[i];
      upd(p[i], gw, m[i], v[i]);
      if (write_g) g[i] = gw;
    }
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += ADAM_THREADS) {
      float gw = g[i];
      upd(p[i], gw, m[i], v[i]);
      if (write_g) g[i] = gw;
    }
  }
}
}  // namespace
}  // namespace msmc

extern "C" int msmc_adam_chunk_elems(void) { return msmc::ADAM_CHUNK; }

extern "C" int msmc_adam_multi(const uint64_t* table, int32_t n_tensors, const int64_t* sizes,
                               const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                               const int32_t* step_index, float* steps, float* partial, float max_norm,
                               float* norm_out, const float* lr, float beta1, float beta2, float eps,
                               float weight_decay, int32_t decoupled, void* stream) {
  MSMC_REQUIRE(table && sizes && chunk_tensor && chunk_index && step_index && steps && lr && n_tensors > 0 &&
               n_chunks > 0);
  MSMC_REQUIRE(partial != nullptr || max_norm <= 0.f);
  cudaStream_t st = (cudaStream_t)stream;
  const bool clip = partial != nullptr && max_norm > 0.f;
  msmc::adam_prepare_kernel<<<n_chunks, msmc::ADAM_THREADS, 0, st>>>(
      reinterpret_cast<const unsigned long long*>(table), n_tensors, reinterpret_cast<const long long*>(sizes),
      chunk_tensor, chunk_index, step_index, steps, partial, clip ? 1 : 0);
  msmc::adam_multi_kernel<<<n_chunks, msmc::ADAM_THREADS, 0, st>>>(
      reinterpret_cast<const unsigned long long*>(table), n_tensors, reinterpret_cast<const long long*>(sizes),
      chunk_tensor, chunk_index, step_index, steps, clip ? partial : nullptr, n_chunks, max_norm, norm_out, lr,
      beta1, beta2, eps, weight_decay, decoupled);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

// ------------------------------------------------------------------------------------------------
// Feature-matching loss over a LIST of tensor pairs in one launch (+ one tiny reduce):
//   loss = sum_t mean |a_t - b_t|          (reference trainers/msmctts_trainer.py:186-190: 55 F.l1_loss terms)
// torch spends ~7 kernels per pair (sub, abs, mean; sign, mul, mul, add in backward) = ~400 launches per step;
// this is 2 + 1.  Block = one 16k-element chunk of one pair; per-block partials are reduced in a fixed order by a
// single block, so the result is run-to-run deterministic.  Backward: ga_t = g * sign(a_t - b_t) / n_t.
// ------------------------------------------------------------------------------------------------
namespace msmc {
namespace {
constexpr int L1_CHUNK = 16384;
constexpr int L1_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (L1_THREADS / 32) ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

__global__ void __launch_bounds__(L1_THREADS) l1_multi_fwd_kernel(
    const unsigned long long* __restrict__ table, int n_tensors, const long long* __restrict__ sizes,
    const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_index, float* __restrict__ partial) {
  __shared__ float sh[L1_THREADS / 32];
  const int t = chunk_tensor[blockIdx.x];
  const long long beg = (long long)chunk_index[blockIdx.x] * L1_CHUNK;
  const long long n = sizes[t];
  const long long end = beg + L1_CHUNK < n ? beg + L1_CHUNK : n;
  const float* __restrict__ a = reinterpret_cast<const float*>(table[t]);
  const float* __restrict__ b = reinterpret_cast<const float*>(table[n_tensors + t]);
  float acc = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  if (vec) {
    const long long end4 = beg + ((end - beg) & ~3LL);
    for (long long i = beg + 4LL * threadIdx.x; i < end4; i += 4LL * L1_THREADS) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + i));
      const float4 y = __ldg(reinterpret_cast<const float4*>(b + i));
      acc += (fabsf(x.x - y.x) + fabsf(x.y - y.y)) + (fabsf(x.z - y.z) + fabsf(x.w - y.w));
    }
    for (long long i = end4 + threadIdx.x; i < end; i += L1_THREADS) acc += fabsf(a[i] - b[i]);
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += L1_THREADS) acc += fabsf(a[i] - b[i]);
  }
  const float tot = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot / (float)n;
}

__global__ void __launch_bounds__(L1_THREADS) l1_multi_reduce_kernel(const float* __restrict__ partial, int n_chunks,
                                                                     float* __restrict__ out) {
  __shared__ float sh[L1_THREADS / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_chunks; i += L1_THREADS) acc += partial[i];
  const float tot = block_sum_256(acc, sh);
  if (threadIdx.x == 0) out[0] = tot;
}

__global__ void __launch_bounds__(L1_THREADS) l1_multi_bwd_kernel(
    const unsigned long long* __restrict__ table, int n_tensors, const long long* __restrict__ sizes,
    const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_index, const float* __restrict__ gout) {
  const int t = chunk_tensor[blockIdx.x];
  const long long beg = (long long)chunk_index[blockIdx.x] * L1_CHUNK;
  const long long n = sizes[t];
  const long long end = beg + L1_CHUNK < n ? beg + L1_CHUNK : n;
  const float* __restrict__ a = reinterpret_cast<const float*>(table[t]);
  const float* __restrict__ b = reinterpret_cast<const float*>(table[n_tensors + t]);
  float* __restrict__ ga = reinterpret_cast<float*>(table[2 * n_tensors + t]);
  const float s = gout[0] / (float)n;
  auto sg = [s](float x, float y) { const float d = x - y; return d > 0.f ? s : (d < 0.f ? -s : 0.f); };
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(ga)) & 15) == 0;
  if (vec) {
    const long long end4 = beg + ((end - beg) & ~3LL);
    for (long long i = beg + 4LL * threadIdx.x; i < end4; i += 4LL * L1_THREADS) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + i));
      const float4 y = __ldg(reinterpret_cast<const float4*>(b + i));
      *reinterpret_cast<float4*>(ga + i) = make_float4(sg(x.x, y.x), sg(x.y, y.y), sg(x.z, y.z), sg(x.w, y.w));
    }
    for (long long i = end4 + threadIdx.x; i < end; i += L1_THREADS) ga[i] = sg(a[i], b[i]);
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += L1_THREADS) ga[i] = sg(a[i], b[i]);
  }
}
}  // namespace
}  // namespace msmc

extern "C" int msmc_l1_chunk_elems(void) { return msmc::L1_CHUNK; }

extern "C" int msmc_l1_multi_fwd(const uint64_t* table, int32_t n_tensors, const int64_t* sizes,
                                 const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                                 float* partial, float* out, void* stream) {
  MSMC_REQUIRE(table && sizes && chunk_tensor && chunk_index && partial && out && n_tensors > 0 && n_chunks > 0);
  cudaStream_t st = (cudaStream_t)stream;
  msmc::l1_multi_fwd_kernel<<<n_chunks, msmc::L1_THREADS, 0, st>>>(
      reinterpret_cast<const unsigned long long*>(table), n_tensors, reinterpret_cast<const long long*>(sizes),
      chunk_tensor, chunk_index, partial);
  msmc::l1_multi_reduce_kernel<<<1, msmc::L1_THREADS, 0, st>>>(partial, n_chunks, out);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_l1_multi_bwd(const uint64_t* table, int32_t n_tensors, const int64_t* sizes,
                                 const int32_t* chunk_tensor, const int32_t* chunk_index, int32_t n_chunks,
                                 const float* gout, void* stream) {
  MSMC_REQUIRE(table && sizes && chunk_tensor && chunk_index && gout && n_tensors > 0 && n_chunks > 0);
  msmc::l1_multi_bwd_kernel<<<n_chunks, msmc::L1_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(table), n_tensors, reinterpret_cast<const long long*>(sizes),
      chunk_tensor, chunk_index, gout);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
