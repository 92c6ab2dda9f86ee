// HBM-bound row / point-wise kernels of the hot path:
//   * dropout + residual + LayerNorm + pad mask (transformer.py:275-283, 374-380, 201-204), forward and backward
//   * STFT magnitude, 'double'-domain mel compression, MelLoss log compression (utils/audio.py:403-419,
//     criterions/stft_loss.py:103-114), forward and backward
//   * gated tanh*sigmoid activation of the ResStack (vqgantts/modules.py:172-179), forward and backward
// All are one pass over their operands with 128-bit accesses where the layout allows.
#include "common.cuh"
#include <algorithm>

namespace msmc {
namespace {

constexpr int LN_MAX_PER_LANE = 32;  // C <= 1024

__device__ __forceinline__ float drop_scale(float p, uint64_t seed, uint64_t salt, uint64_t index) {
  if (p <= 0.f) return 1.f;
  return uniform01(seed, salt, index) >= p ? 1.f / (1.f - p) : 0.f;
}

// one warp per row; PER = register slots per lane (>= ceil(C/32))
template <int PER>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ r,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         const int* __restrict__ lengths, float* __restrict__ y, float* __restrict__ xhat,
                         float* __restrict__ rstd, int rows, int t, int C, float eps, float drop_p,
                         const uint64_t* __restrict__ seed_ptr, uint64_t salt) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const uint64_t seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps_per_grid) {
    float x[PER];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      x[j] = 0.f;
      {
        const int c = lane + 32 * j;
        if (c < C) {
          const int64_t e = (int64_t)row * C + c;
          float v = a[e] * drop_scale(drop_p, seed, salt, (uint64_t)e);
          if (r) v += r[e];
          x[j] = v;
          s += v;
        }
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j)
      {
        const int c = lane + 32 * j;
        if (c < C) { const float d = x[j] - mean; sq = fmaf(d, d, sq); }
      }
    const float var = warp_sum(sq) / (float)C;
    const float rs = rsqrtf(var + eps);
    const int b = row / t, i = row - b * t;
    const float mk = (lengths == nullptr || i < lengths[b]) ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j)
      {
        const int c = lane + 32 * j;
        if (c < C) {
          const int64_t e = (int64_t)row * C + c;
          const float xh = (x[j] - mean) * rs;
          xhat[e] = xh;
          y[e] = mk * fmaf(xh, gamma[c], beta[c]);
        }
      }
    if (lane == 0) rstd[row] = rs;
  }
}

// backward: per-row input gradient + per-CTA partial dgamma/dbeta
template <int PER>
__global__ void __launch_bounds__(256)
add_layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ xhat,
                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                         const int* __restrict__ lengths, float* __restrict__ ga, float* __restrict__ gr,
                         float* __restrict__ partial, int rows, int t, int C, float drop_p,
                         const uint64_t* __restrict__ seed_ptr, uint64_t salt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const uint64_t seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  float dg[PER], db[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) { dg[j] = 0.f; db[j] = 0.f; }
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps_per_grid) {
    const int b = row / t, i = row - b * t;
    const float mk = (lengths == nullptr || i < lengths[b]) ? 1.f : 0.f;
    const float rs = rstd[row];
    float g[PER], xh[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      g[j] = 0.f; xh[j] = 0.f;
      {
        const int c = lane + 32 * j;
        if (c < C) {
          const int64_t e = (int64_t)row * C + c;
          const float gv = gy[e] * mk;
          xh[j] = xhat[e];
          dg[j] = fmaf(gv, xh[j], dg[j]);
          db[j] += gv;
          g[j] = gv * gamma[c];
          s1 += g[j];
          s2 = fmaf(g[j], xh[j], s2);
        }
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int j = 0; j < PER; ++j)
      {
        const int c = lane + 32 * j;
        if (c < C) {
          const int64_t e = (int64_t)row * C + c;
          const float dx = rs * (g[j] - s1 - xh[j] * s2);
          if (gr) gr[e] = dx;
          ga[e] = dx * drop_scale(drop_p, seed, salt, (uint64_t)e);
        }
      }
  }
  // cross-warp reduction in fixed warp order, one partial row per CTA
  extern __shared__ float red[];  // [8][2*C]
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int c = lane + 32 * j;
    if (c < C) {
      red[(size_t)warp * 2 * C + c] = dg[j];
      red[(size_t)warp * 2 * C + C + c] = db[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[(size_t)w * 2 * C + c];
    partial[(size_t)blockIdx.x * 2 * C + c] = s;
  }
}

// column sums of the per-CTA partial rows: 32 columns per CTA, 8 warps each summing every 8th partial row (four
// independent loads in flight), then a fixed-order sum over the warps -- deterministic, and ~60 loads per thread
// instead of one thread walking all `nparts` rows of its column (32 us per call on the encoder's backward chain)
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ out0,
                       float* __restrict__ out1, int split) {
  // out0 gets columns [0, split), out1 gets [split, n)
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < n) {
    int p = w;
    for (; p + 24 < nparts; p += 32) {
      s0 += partial[(size_t)p * n + c];
      s1 += partial[(size_t)(p + 8) * n + c];
      s2 += partial[(size_t)(p + 16) * n + c];
      s3 += partial[(size_t)(p + 24) * n + c];
    }
    for (; p < nparts; p += 8) s0 += partial[(size_t)p * n + c];
  }
  red[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && c < n) {
    float s = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][lane];
    if (c < split) out0[c] = s; else out1[c - split] = s;
  }
}

// ---------------------------------------------------------------------------------------------- spectral
__global__ void spec_mag_fwd_kernel(const float* __restrict__ spec, float* __restrict__ mag, int64_t rows,
                                    int F, int Fp, float floor_, int floor_add) {
  const int64_t total = rows * F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / F;
    const int f = (int)(e - r * F);
    const float re = spec[r * 2 * Fp + f], im = spec[r * 2 * Fp + Fp + f];
    const float p = re * re + im * im;
    mag[e] = sqrtf(floor_add ? p + floor_ : fmaxf(p, floor_));
  }
}
__global__ void spec_mag_bwd_kernel(const float* __restrict__ gmag, const float* __restrict__ spec,
                                    const float* __restrict__ mag, float* __restrict__ gspec, int64_t rows,
                                    int F, int Fp, float floor_, int floor_add) {
  // one thread per (row, padded bin): the padding columns of the gradient are written as zeros
  const int64_t total = rows * Fp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / Fp;
    const int f = (int)(e - r * Fp);
    float gre = 0.f, gim = 0.f;
    if (f < F) {
      const float re = spec[r * 2 * Fp + f], im = spec[r * 2 * Fp + Fp + f];
      const float p = re * re + im * im;
      // d sqrt(u)/du = 1/(2 sqrt(u)); clamp passes gradient only where p >= floor
      float c = 0.f;
      if (floor_add || p >= floor_) c = gmag[r * F + f] / mag[r * F + f];
      gre = c * re;
      gim = c * im;
    }
    gspec[r * 2 * Fp + f] = gre;
    gspec[r * 2 * Fp + Fp + f] = gim;
  }
}

__global__ void mel_double_fwd_kernel(const float* __restrict__ mel, float* __restrict__ out, int64_t n,
                                      float ref_db, float min_db) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float m = mel[e];
    float lg = 20.f * log10f(m) - ref_db;
    lg = (lg - min_db) / (-min_db);
    lg = fminf(fmaxf(lg, 0.f), 1.f);
    reinterpret_cast<float2*>(out)[e] = make_float2(m, lg);
  }
}
__global__ void mel_double_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ mel,
                                      float* __restrict__ gmel, int64_t n, float ref_db, float min_db) {
  const float k = 20.f / 2.302585092994046f / (-min_db);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float m = mel[e];
    const float2 g = reinterpret_cast<const float2*>(gout)[e];
    float lg = 20.f * log10f(m) - ref_db;
    lg = (lg - min_db) / (-min_db);
    float gm = g.x;
    if (lg >= 0.f && lg <= 1.f) gm += g.y * k / m;   // clamp passes gradient on the closed interval
    gmel[e] = gm;
  }
}

// out = xf(v, aux): one vectorised pass.  Used by the conv backward to materialise the pre-activation gradient
// g * act'(y) ONCE, so that the data-gradient and the weight-gradient kernels both read a plain operand (their
// aux-reading producer variants were 2-4x slower than the plain ones: profiles/r01 launch lists).
__global__ void __launch_bounds__(256) xform_apply_kernel(const float* __restrict__ v, const float* __restrict__ aux,
                                                          float* __restrict__ out, int64_t n, int xf, float slope) {
  const int64_t n4 = n >> 2;
  const bool need_aux = xf_needs_aux(xf);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(v) + e);
    const float4 a = need_aux ? __ldg(reinterpret_cast<const float4*>(aux) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(out)[e] = make_float4(apply_xf(xf, slope, x.x, a.x), apply_xf(xf, slope, x.y, a.y),
                                                    apply_xf(xf, slope, x.z, a.z), apply_xf(xf, slope, x.w, a.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t e = (n4 << 2) + threadIdx.x;
    out[e] = apply_xf(xf, slope, v[e], need_aux ? aux[e] : 0.f);
  }
}

__global__ void log_clamp_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float clip) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    y[e] = logf(fmaxf(x[e], clip));
}
__global__ void log_clamp_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                     float* __restrict__ gx, int64_t n, float clip) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[e];
    gx[e] = v >= clip ? gy[e] / v : 0.f;
  }
}

__global__ void gated_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / C;
    const int c = (int)(e - r * C);
    const float a = x[r * 2 * C + c], b = x[r * 2 * C + C + c];
    y[e] = tanhf(a) * (1.f / (1.f + expf(-b)));
  }
}
__global__ void gated_act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                     float* __restrict__ gx, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / C;
    const int c = (int)(e - r * C);
    const float a = x[r * 2 * C + c], b = x[r * 2 * C + C + c];
    const float th = tanhf(a), sg = 1.f / (1.f + expf(-b));
    const float g = gy[e];
    gx[r * 2 * C + c] = g * sg * (1.f - th * th);
    gx[r * 2 * C + C + c] = g * th * sg * (1.f - sg);
  }
}

// Backward of reflect-padded STFT framing: per-frame time-domain gradients (B, frames, win) are overlap-added onto
// the padded time axis and the reflected borders folded back in one pass:
//   gx[b, l] = sum_{p in preimage(l)} sum_{f : 0 <= p - f*hop < win} gframes[b, f, p - f*hop]
__global__ void overlap_add_fold_kernel(const float* __restrict__ gf, float* __restrict__ gx, int B, int frames,
                                        int win, int hop, int L, int pad) {
  const int64_t total = (int64_t)B * L;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / L), l = (int)(e - (int64_t)b * L);
    int pc[3], np = 0;
    pc[np++] = l + pad;
    if (l >= 1 && l <= pad) pc[np++] = pad - l;
    if (l <= L - 2 && l >= L - 1 - pad) pc[np++] = pad + 2 * (L - 1) - l;
    float s = 0.f;
    for (int i = 0; i < np; ++i) {
      const int p = pc[i];
      int f_hi = p / hop;                       // largest f with f*hop <= p
      if (f_hi > frames - 1) f_hi = frames - 1;
      for (int f = f_hi; f >= 0 && p - f * hop < win; --f)
        s += gf[((int64_t)b * frames + f) * win + (p - f * hop)];
    }
    gx[e] = s;
  }
}

// STFT framing: frames[b*F + f][k] = x[b][reflect(f*hop + k - pad)] for k < win, 0 for win <= k < win_p.
// Materialising the (B*frames, win_p) matrix (a few MB) turns the windowed DFT into one tensor-core GEMM.
__global__ void frame_unfold_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int L, int frames,
                                    int win, int win_p, int hop, int pad) {
  const int64_t total = (int64_t)B * frames * win_p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % win_p);
    const int64_t r = e / win_p;
    const int f = (int)(r % frames), b = (int)(r / frames);
    float v = 0.f;
    if (k < win) {
      int i = f * hop + k - pad;
      if (i < 0) i = -i;
      if (i >= L) i = 2 * (L - 1) - i;
      v = x[(int64_t)b * L + i];
    }
    out[e] = v;
  }
}

inline int ew_blocks(int64_t n) { return (int)std::min<int64_t>(ceil_div64(n, 256), (int64_t)num_sms() * 16); }

constexpr int LN_BWD_MAX_BLOCKS = 296;

}  // namespace
}  // namespace msmc

using namespace msmc;

extern "C" int msmc_add_layernorm_fwd(const float* a, const float* r, const float* gamma, const float* beta,
                                      const int32_t* lengths, float* y, float* xhat, float* rstd, int32_t B,
                                      int32_t t, int32_t C, float eps, float drop_p, const uint64_t* seed,
                                      uint64_t call_salt, void* stream) {
  MSMC_REQUIRE(a && gamma && beta && y && xhat && rstd && B > 0 && t > 0 && C > 0);
  MSMC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed));
  if (C > 32 * LN_MAX_PER_LANE) return MSMC_ERR_UNSUPPORTED;
  const int rows = B * t;
  const int blocks = std::min(ceil_div(rows, 8), num_sms() * 8);
  const int per = ceil_div(C, 32);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_FWD(P) add_layernorm_fwd_kernel<P><<<blocks, 256, 0, st>>>(a, r, gamma, beta, lengths, y, xhat, rstd, rows, t, C, eps, drop_p, seed, call_salt)
  if (per <= 2) LN_FWD(2); else if (per <= 4) LN_FWD(4); else if (per <= 8) LN_FWD(8);
  else if (per <= 16) LN_FWD(16); else if (per <= 20) LN_FWD(20); else LN_FWD(32);
#undef LN_FWD
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int64_t msmc_add_layernorm_bwd_workspace(int32_t C) {
  return (int64_t)LN_BWD_MAX_BLOCKS * 2 * C * (int64_t)sizeof(float);
}

extern "C" int msmc_add_layernorm_bwd(const float* gy, const float* xhat, const float* rstd, const float* gamma,
                                      const int32_t* lengths, float* ga, float* gr, float* dgamma, float* dbeta,
                                      float* workspace, int32_t B, int32_t t, int32_t C, float drop_p,
                                      const uint64_t* seed, uint64_t call_salt, void* stream) {
  MSMC_REQUIRE(gy && xhat && rstd && gamma && ga && dgamma && dbeta && workspace && B > 0 && t > 0 && C > 0);
  MSMC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed));
  if (C > 32 * LN_MAX_PER_LANE) return MSMC_ERR_UNSUPPORTED;
  const int rows = B * t;
  const int blocks = std::min(ceil_div(rows, 8), LN_BWD_MAX_BLOCKS);
  const size_t smem = (size_t)8 * 2 * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const int per = ceil_div(C, 32);
#define LN_BWD(P)                                                                                              \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      cudaFuncSetAttribute(add_layernorm_bwd_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    add_layernorm_bwd_kernel<P><<<blocks, 256, smem, st>>>(gy, xhat, rstd, gamma, lengths, ga, gr, workspace,  \
                                                           rows, t, C, drop_p, seed, call_salt);               \
  } while (0)
  if (per <= 2) LN_BWD(2); else if (per <= 4) LN_BWD(4); else if (per <= 8) LN_BWD(8);
  else if (per <= 16) LN_BWD(16); else if (per <= 20) LN_BWD(20); else LN_BWD(32);
#undef LN_BWD
  MSMC_CHECK_LAUNCH();
  colsum_partials_kernel<<<ceil_div(2 * C, 32), 256, 0, st>>>(workspace, blocks, 2 * C, dgamma, dbeta, C);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_spec_magnitude_fwd(const float* spec, float* mag, int64_t rows, int32_t F, int32_t Fp,
                                       float floor_, int32_t floor_add, void* stream) {
  MSMC_REQUIRE(spec && mag && rows > 0 && F > 0 && Fp >= F);
  spec_mag_fwd_kernel<<<ew_blocks(rows * F), 256, 0, (cudaStream_t)stream>>>(spec, mag, rows, F, Fp, floor_,
                                                                             floor_add);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_spec_magnitude_bwd(const float* gmag, const float* spec, const float* mag, float* gspec,
                                       int64_t rows, int32_t F, int32_t Fp, float floor_, int32_t floor_add,
                                       void* stream) {
  MSMC_REQUIRE(gmag && spec && mag && gspec && rows > 0 && F > 0 && Fp >= F);
  spec_mag_bwd_kernel<<<ew_blocks(rows * Fp), 256, 0, (cudaStream_t)stream>>>(gmag, spec, mag, gspec, rows, F, Fp,
                                                                              floor_, floor_add);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_mel_double_fwd(const float* mel, float* out, int64_t n, float ref_db, float min_db,
                                   void* stream) {
  MSMC_REQUIRE(mel && out && n > 0 && min_db < 0.f);
  mel_double_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(mel, out, n, ref_db, min_db);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_mel_double_bwd(const float* gout, const float* mel, float* gmel, int64_t n, float ref_db,
                                   float min_db, void* stream) {
  MSMC_REQUIRE(gout && mel && gmel && n > 0 && min_db < 0.f);
  mel_double_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(gout, mel, gmel, n, ref_db, min_db);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_xform_apply(const float* v, const float* aux, float* out, int64_t n, int32_t xf, float slope,
                                void* stream) {
  MSMC_REQUIRE(v && out && n > 0 && xf >= MSMC_XF_NONE && xf <= MSMC_XF_MUL_DTANH);
  MSMC_REQUIRE(!xf_needs_aux(xf) || aux);
  MSMC_REQUIRE(((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(aux)) & 15) == 0);
  xform_apply_kernel<<<ew_blocks((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(v, aux, out, n, xf, slope);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_log_clamp_fwd(const float* x, float* y, int64_t n, float clip, void* stream) {
  MSMC_REQUIRE(x && y && n > 0);
  log_clamp_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, clip);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_log_clamp_bwd(const float* gy, const float* x, float* gx, int64_t n, float clip,
                                  void* stream) {
  MSMC_REQUIRE(gy && x && gx && n > 0);
  log_clamp_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(gy, x, gx, n, clip);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_overlap_add_fold(const float* gframes, float* gx, int32_t B, int32_t frames, int32_t win,
                                     int32_t hop, int32_t L, int32_t pad, void* stream) {
  MSMC_REQUIRE(gframes && gx && B > 0 && frames > 0 && win > 0 && hop > 0 && L > 0 && pad >= 0 && pad < L);
  overlap_add_fold_kernel<<<ew_blocks((int64_t)B * L), 256, 0, (cudaStream_t)stream>>>(gframes, gx, B, frames, win,
                                                                                      hop, L, pad);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_frame_unfold(const float* x, float* frames_out, int32_t B, int32_t L, int32_t frames, int32_t win,
                                 int32_t win_p, int32_t hop, int32_t pad, void* stream) {
  MSMC_REQUIRE(x && frames_out && B > 0 && L > 1 && frames > 0 && win > 0 && win_p >= win && hop > 0 && pad >= 0 &&
               pad < L);
  frame_unfold_kernel<<<ew_blocks((int64_t)B * frames * win_p), 256, 0, (cudaStream_t)stream>>>(
      x, frames_out, B, L, frames, win, win_p, hop, pad);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_gated_act_fwd(const float* x, float* y, int64_t rows, int32_t C, void* stream) {
  MSMC_REQUIRE(x && y && rows > 0 && C > 0);
  gated_act_fwd_kernel<<<ew_blocks(rows * C), 256, 0, (cudaStream_t)stream>>>(x, y, rows, C);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
extern "C" int msmc_gated_act_bwd(const float* gy, const float* x, float* gx, int64_t rows, int32_t C,
                                  void* stream) {
  MSMC_REQUIRE(gy && x && gx && rows > 0 && C > 0);
  gated_act_bwd_kernel<<<ew_blocks(rows * C), 256, 0, (cudaStream_t)stream>>>(gy, x, gx, rows, C);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
