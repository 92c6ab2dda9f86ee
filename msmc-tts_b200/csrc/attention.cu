// FFT-block self-attention, forward and backward, fp32 CUDA-core tiles with online softmax (no (t,t) matrix in
// HBM).  Replaces MultiHeadAttention's permute-copies + ScaledDotProductAttention (bmm, /temperature,
// masked_fill(-inf) on padded keys, softmax(dim=2), dropout, bmm) -- reference
// acoustic_models/transformer.py:246-328.  Head dim is 64 (every in-tree config).
//
// qkv : (B, t, n_head*3*D) straight out of the fused QKV Linear; per head h the columns are
//       [h*3D, h*3D+D) = q, [.. +D, +2D) = k, [.. +2D, +3D) = v   (transformer.py:253-260)
// out : (B, t, n_head*D) with head-major columns (transformer.py:270-272)
#include "common.cuh"
#include <cstdlib>

namespace msmc {
namespace {

constexpr int D = 64;      // head dim
constexpr int TQ = 64;     // query tile
constexpr int TK = 64;     // key tile
constexpr int PITCH = 68;  // smem row pitch (floats), 16B-aligned rows
constexpr int TILE = 64 * PITCH;

// global (rows x D, row pitch ld) -> smem transposed [d][row]
__device__ __forceinline__ void load_tile_T(float* s, const float* g, int64_t ld, int row0, int nrows_total) {
  for (int e = threadIdx.x; e < 64 * D; e += blockDim.x) {
    const int r = e >> 6, d = e & 63;
    const int gr = row0 + r;
    s[d * PITCH + r] = (gr < nrows_total) ? __ldg(g + (int64_t)gr * ld + d) : 0.f;
  }
}
// global -> smem natural [row][d]
__device__ __forceinline__ void load_tile_N(float* s, const float* g, int64_t ld, int row0, int nrows_total) {
  for (int e = threadIdx.x; e < 64 * D; e += blockDim.x) {
    const int r = e >> 6, d = e & 63;
    const int gr = row0 + r;
    s[r * PITCH + d] = (gr < nrows_total) ? __ldg(g + (int64_t)gr * ld + d) : 0.f;
  }
}

// c[i][j] += sum_k At[k][r0+i] * Bt[k][c0+j]   (both operands stored [k][index])
__device__ __forceinline__ void mma_TT(float (&c)[4][4], const float* At, const float* Bt, int r0, int c0) {
#pragma unroll 8
  for (int k = 0; k < 64; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(At + k * PITCH + r0);
    const float4 b = *reinterpret_cast<const float4*>(Bt + k * PITCH + c0);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = fmaf(av[i], bv[j], c[i][j]);
  }
}
// c[i][j] += sum_k A[r0+i][k] * B[k][c0+j]    (A natural [row][k], B natural [k][col])
__device__ __forceinline__ void mma_NN(float (&c)[4][4], const float* A, const float* B, int r0, int c0) {
#pragma unroll 8
  for (int k = 0; k < 64; ++k) {
    float av[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = A[(r0 + i) * PITCH + k];
    const float4 b = *reinterpret_cast<const float4*>(B + k * PITCH + c0);
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = fmaf(av[i], bv[j], c[i][j]);
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float keep_scale(float drop_p, uint64_t seed, uint64_t salt, int bh, int t, int q, int k) {
  if (drop_p <= 0.f) return 1.f;
  const uint64_t index = ((uint64_t)bh * (uint64_t)t + (uint64_t)q) * (uint64_t)t + (uint64_t)k;
  return uniform01(seed, salt, index) >= drop_p ? 1.f / (1.f - drop_p) : 0.f;
}

__global__ void __launch_bounds__(256)
attention_fwd_kernel(const float* __restrict__ qkv, const int* __restrict__ lengths, float* __restrict__ out,
                     float* __restrict__ lse, int B, int t, int n_head, float inv_temp, float drop_p,
                     const uint64_t* __restrict__ seed_ptr, uint64_t salt) {
  extern __shared__ __align__(16) float sm[];
  float* Qt = sm;
  float* Kt = sm + TILE;
  float* Vs = sm + 2 * TILE;
  float* Ps = sm + 3 * TILE;
  const int bh = blockIdx.y, h = bh / B, b = bh % B;  // reference batches heads as (n_head*B): index = h*B + b
  const int q0 = blockIdx.x * TQ;
  const int64_t ld = (int64_t)n_head * 3 * D;
  const float* base = qkv + (int64_t)b * t * ld + h * 3 * D;
  const int len = lengths[b];
  const uint64_t seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int r0 = ty * 4, c0 = tx * 4;

  load_tile_T(Qt, base, ld, q0, t);
  float o[4][4], m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < t; k0 += TK) {
    __syncthreads();
    load_tile_T(Kt, base + D, ld, k0, t);
    load_tile_N(Vs, base + 2 * D, ld, k0, t);
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    mma_TT(s, Qt, Kt, r0, c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float rmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + c0 + j;
        s[i][j] = (key < len && key < t) ? s[i][j] * inv_temp : -INFINITY;
        rmax = fmaxf(rmax, s[i][j]);
      }
      rmax = group16_max(rmax);
      const float m_new = fmaxf(m[i], rmax);
      const float corr = (m_new == -INFINITY) ? 1.f : __expf(m[i] - m_new);
      float rsum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (m_new == -INFINITY) ? 0.f : __expf(s[i][j] - m_new);
        rsum += p;
        const float ks = keep_scale(drop_p, seed, salt, bh, t, q0 + r0 + i, k0 + c0 + j);
        Ps[(r0 + i) * PITCH + c0 + j] = p * ks;
      }
      rsum = group16_sum(rsum);
      l[i] = l[i] * corr + rsum;
      m[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
    }
    __syncthreads();
    mma_NN(o, Ps, Vs, r0, c0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + r0 + i;
    if (q < t) {
      const float inv_l = l[i] > 0.f ? 1.f / l[i] : 0.f;
      float4 v = make_float4(o[i][0] * inv_l, o[i][1] * inv_l, o[i][2] * inv_l, o[i][3] * inv_l);
      *reinterpret_cast<float4*>(out + ((int64_t)b * t + q) * (n_head * D) + h * D + c0) = v;
      if (tx == 0) lse[(int64_t)bh * t + q] = m[i] + logf(l[i]);
    }
  }
}

// dQ: one CTA per (query tile, b, h), loops over key tiles
__global__ void __launch_bounds__(256)
attention_bwd_dq_kernel(const float* __restrict__ qkv, const int* __restrict__ lengths,
                        const float* __restrict__ out, const float* __restrict__ lse,
                        const float* __restrict__ gout, float* __restrict__ gqkv, int B, int t, int n_head,
                        float inv_temp, float drop_p, const uint64_t* __restrict__ seed_ptr, uint64_t salt) {
  extern __shared__ __align__(16) float sm[];
  float* Qt = sm;
  float* dOt = sm + TILE;
  float* Kt = sm + 2 * TILE;
  float* Ks = sm + 3 * TILE;
  float* Vt = sm + 4 * TILE;
  float* Ps = sm + 5 * TILE;
  const int bh = blockIdx.y, h = bh / B, b = bh % B;
  const int q0 = blockIdx.x * TQ;
  const int64_t ld = (int64_t)n_head * 3 * D;
  const int64_t ldo = (int64_t)n_head * D;
  const float* base = qkv + (int64_t)b * t * ld + h * 3 * D;
  const float* go = gout + (int64_t)b * t * ldo + h * D;
  const float* oo = out + (int64_t)b * t * ldo + h * D;
  const int len = lengths[b];
  const uint64_t seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int r0 = ty * 4, c0 = tx * 4;

  load_tile_T(Qt, base, ld, q0, t);
  load_tile_T(dOt, go, ldo, q0, t);
  float Dr[4], L[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + r0 + i;
    float s = 0.f;
    if (q < t) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(go + (int64_t)q * ldo + c0));
      const float4 c = __ldg(reinterpret_cast<const float4*>(oo + (int64_t)q * ldo + c0));
      s = a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    }
    Dr[i] = group16_sum(s);
    L[i] = (q < t) ? lse[(int64_t)bh * t + q] : 0.f;
  }
  float dq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

  for (int k0 = 0; k0 < t; k0 += TK) {
    __syncthreads();
    load_tile_T(Kt, base + D, ld, k0, t);
    load_tile_N(Ks, base + D, ld, k0, t);
    load_tile_T(Vt, base + 2 * D, ld, k0, t);
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
    mma_TT(s, Qt, Kt, r0, c0);
    mma_TT(dp, dOt, Vt, r0, c0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = k0 + c0 + j;
        float ds = 0.f;
        if (key < len && key < t) {
          const float p = __expf(s[i][j] * inv_temp - L[i]);
          const float ks = keep_scale(drop_p, seed, salt, bh, t, q0 + r0 + i, key);
          ds = p * (dp[i][j] * ks - Dr[i]) * inv_temp;
        }
        Ps[(r0 + i) * PITCH + c0 + j] = ds;
      }
    __syncthreads();
    mma_NN(dq, Ps, Ks, r0, c0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + r0 + i;
    if (q < t)
      *reinterpret_cast<float4*>(gqkv + ((int64_t)b * t + q) * ld + h * 3 * D + c0) =
          make_float4(dq[i][0], dq[i][1], dq[i][2], dq[i][3]);
  }
}

// dK, dV: one CTA per (key tile, b, h), loops over query tiles
__global__ void __launch_bounds__(256)
attention_bwd_dkv_kernel(const float* __restrict__ qkv, const int* __restrict__ lengths,
                         const float* __restrict__ out, const float* __restrict__ lse,
                         const float* __restrict__ gout, float* __restrict__ gqkv, int B, int t, int n_head,
                         float inv_temp, float drop_p, const uint64_t* __restrict__ seed_ptr, uint64_t salt) {
  extern __shared__ __align__(16) float sm[];
  float* Kt = sm;
  float* Vt = sm + TILE;
  float* Qt = sm + 2 * TILE;
  float* Qs = sm + 3 * TILE;
  float* dOt = sm + 4 * TILE;
  float* dOs = sm + 5 * TILE;
  float* Pt = sm + 6 * TILE;   // [key][query]  P_drop^T
  float* dSt = sm + 7 * TILE;  // [key][query]  dS^T
  float* Dq = sm + 8 * TILE;   // [64] D per query, then [64] lse
  const int bh = blockIdx.y, h = bh / B, b = bh % B;
  const int k0 = blockIdx.x * TK;
  const int64_t ld = (int64_t)n_head * 3 * D;
  const int64_t ldo = (int64_t)n_head * D;
  const float* base = qkv + (int64_t)b * t * ld + h * 3 * D;
  const float* go = gout + (int64_t)b * t * ldo + h * D;
  const float* oo = out + (int64_t)b * t * ldo + h * D;
  const int len = lengths[b];
  const uint64_t seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int r0 = ty * 4, c0 = tx * 4;   // r0: keys, c0: queries (for S^T) or d (for dK/dV)

  load_tile_T(Kt, base + D, ld, k0, t);
  load_tile_T(Vt, base + 2 * D, ld, k0, t);
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }

  for (int q0 = 0; q0 < t; q0 += TQ) {
    __syncthreads();
    load_tile_T(Qt, base, ld, q0, t);
    load_tile_N(Qs, base, ld, q0, t);
    load_tile_T(dOt, go, ldo, q0, t);
    load_tile_N(dOs, go, ldo, q0, t);
    if (threadIdx.x < 64) {
      const int q = q0 + threadIdx.x;
      float s = 0.f, lv = 0.f;
      if (q < t) {
        for (int d = 0; d < D; ++d) s = fmaf(__ldg(go + (int64_t)q * ldo + d), __ldg(oo + (int64_t)q * ldo + d), s);
        lv = lse[(int64_t)bh * t + q];
      }
      Dq[threadIdx.x] = s;
      Dq[64 + threadIdx.x] = lv;
    }
    __syncthreads();
    float st[4][4], dpt[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { st[i][j] = 0.f; dpt[i][j] = 0.f; }
    mma_TT(st, Kt, Qt, r0, c0);     // S^T[key][q]
    mma_TT(dpt, Vt, dOt, r0, c0);   // dP_drop^T[key][q]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int key = k0 + r0 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + c0 + j;
        float pd = 0.f, ds = 0.f;
        if (key < len && key < t && q < t) {
          const float p = __expf(st[i][j] * inv_temp - Dq[64 + c0 + j]);
          const float ks = keep_scale(drop_p, seed, salt, bh, t, q, key);
          pd = p * ks;
          ds = p * (dpt[i][j] * ks - Dq[c0 + j]) * inv_temp;
        }
        Pt[(r0 + i) * PITCH + c0 + j] = pd;
        dSt[(r0 + i) * PITCH + c0 + j] = ds;
      }
    }
    __syncthreads();
    mma_NN(dv, Pt, dOs, r0, c0);
    mma_NN(dk, dSt, Qs, r0, c0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = k0 + r0 + i;
    if (key < t) {
      float* row = gqkv + ((int64_t)b * t + key) * ld + h * 3 * D;
      *reinterpret_cast<float4*>(row + D + c0) = make_float4(dk[i][0], dk[i][1], dk[i][2], dk[i][3]);
      *reinterpret_cast<float4*>(row + 2 * D + c0) = make_float4(dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
    }
  }
}

}  // namespace
}  // namespace msmc

using namespace msmc;

extern "C" int msmc_attention_fwd(const float* qkv, const int32_t* lengths, float* out, float* lse, int32_t B,
                                  int32_t t, int32_t n_head, int32_t d, float inv_temperature, float drop_p,
                                  const uint64_t* seed, uint64_t call_salt, void* stream) {
  MSMC_REQUIRE(qkv && lengths && out && lse && B > 0 && t > 0 && n_head > 0);
  MSMC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed));
  if (d != D) return MSMC_ERR_UNSUPPORTED;
  const size_t smem = 4 * TILE * sizeof(float);
  cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ceil_div(t, TQ), B * n_head);
  attention_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(qkv, lengths, out, lse, B, t, n_head,
                                                                  inv_temperature, drop_p, seed, call_salt);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_attention_bwd(const float* qkv, const int32_t* lengths, const float* out, const float* lse,
                                  const float* gout, float* gqkv, int32_t B, int32_t t, int32_t n_head,
                                  int32_t d, float inv_temperature, float drop_p, const uint64_t* seed,
                                  uint64_t call_salt, void* stream) {
  MSMC_REQUIRE(qkv && lengths && out && lse && gout && gqkv && B > 0 && t > 0 && n_head > 0);
  MSMC_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed));
  if (d != D) return MSMC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_q = 6 * TILE * sizeof(float);
  const size_t smem_kv = (8 * TILE + 128) * sizeof(float);
  cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
  cudaFuncSetAttribute(attention_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv);
  dim3 gq(ceil_div(t, TQ), B * n_head), gk(ceil_div(t, TK), B * n_head);
  // dQ and dK/dV are independent (they write disjoint columns of gqkv) and neither fills the GPU evenly on its own
  // (256 CTAs: two uneven waves at one CTA per SM for dK/dV): dQ is forked onto a helper stream and joined again, so
  // the block scheduler packs the CTAs of both.  Event fork / join is capturable in CUDA graphs.  One helper stream
  // and event pair per calling thread (the trainer drives the encoder / decoder chain from one thread, one stream).
  struct Fork { cudaStream_t s2 = nullptr; cudaEvent_t e1 = nullptr, e2 = nullptr; int dev = -1; };
  static thread_local Fork fk;
  static const bool use_fork = [] { const char* e = getenv("MSMC_ATTN_BWD_FORK"); return e ? atoi(e) != 0 : false; }();   // (off until measured)
  int dev = 0;
  cudaGetDevice(&dev);
  if (use_fork && fk.dev != dev) {
    fk = Fork();
    if (cudaStreamCreateWithFlags(&fk.s2, cudaStreamNonBlocking) == cudaSuccess &&
        cudaEventCreateWithFlags(&fk.e1, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&fk.e2, cudaEventDisableTiming) == cudaSuccess)
      fk.dev = dev;
    else
      cudaGetLastError();
  }
  const bool fork = use_fork && fk.dev == dev;
  cudaStream_t sq = st;
  if (fork) {
    if (cudaEventRecord(fk.e1, st) != cudaSuccess || cudaStreamWaitEvent(fk.s2, fk.e1, 0) != cudaSuccess)
      return MSMC_ERR_LAUNCH;
    sq = fk.s2;
  }
  attention_bwd_dkv_kernel<<<gk, 256, smem_kv, st>>>(qkv, lengths, out, lse, gout, gqkv, B, t, n_head,
                                                     inv_temperature, drop_p, seed, call_salt);
  MSMC_CHECK_LAUNCH();
  attention_bwd_dq_kernel<<<gq, 256, smem_q, sq>>>(qkv, lengths, out, lse, gout, gqkv, B, t, n_head,
                                                   inv_temperature, drop_p, seed, call_salt);
  MSMC_CHECK_LAUNCH();
  if (fork) {
    if (cudaEventRecord(fk.e2, fk.s2) != cudaSuccess || cudaStreamWaitEvent(st, fk.e2, 0) != cudaSuccess)
      return MSMC_ERR_LAUNCH;
  }
  return MSMC_OK;
}
