// Implicit-GEMM convolution family, fp32 CUDA-core path (first correct path; the tcgen05 TF32 path in
// conv_umma.cu takes the large stride-1 shapes).  One tile engine, three loaders:
//   conv_gemm_kernel  : forward form and conv-transpose form (phase-decomposed so no tap is wasted)
//   conv_wgrad_kernel : weight gradient, split over output rows, partials reduced deterministically
// Replaces torch.nn.Conv1d / ConvTranspose1d / Conv2d / Linear forward+backward at the call sites
// listed in include/msmc_b200.h.
#include "common.cuh"
#include <algorithm>

namespace msmc {
namespace {

constexpr int BM = 128;   // rows of the output tile held by one CTA
constexpr int BK = 16;    // reduction chunk
constexpr int NTHREADS = 256;
constexpr int MAX_TAPS_TABLE = 4096;

struct ConvArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* w;
  const float* bias;
  const float* residual;
  const float* dst_aux;
  float* dst;
  int vec_src;      // 8-wide contiguous channel loads allowed
  int b_kmajor;     // weight tile: lanes run along the reduction index (ws_cs == 1)
};

struct Taps {
  // valid taps of the current phase; forward form decodes directly
  int n;
  const short* kh;
  const short* kw;
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // ReflectionPad semantics: -1 -> 1, n -> n-2 (single bounce is enough for pad < n)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Source coordinate for destination (hd, wd) and tap (kh, kw). Returns false when the tap reads padding zeros.
__device__ __forceinline__ bool src_coord(const msmc_conv_geom& g, int hd, int wd, int kh, int kw, int& hs, int& ws) {
  if (!g.transposed) {
    hs = hd * g.sh + kh * g.dh - g.ph;
    ws = wd * g.sw + kw * g.dw - g.pw;
    if (g.pad_reflect) {
      hs = reflect_idx(hs, g.Hs);
      ws = reflect_idx(ws, g.Ws);
      return true;
    }
    return hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
  } else {
    int th = hd + g.ph - kh * g.dh;
    int tw = wd + g.pw - kw * g.dw;
    if (th < 0 || tw < 0) return false;
    if (th % g.sh != 0 || tw % g.sw != 0) return false;
    hs = th / g.sh;
    ws = tw / g.sw;
    return hs < g.Hs && ws < g.Ws;
  }
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_kernel(const ConvArgs a) {
  constexpr int TN = BN / 16;
  const msmc_conv_geom& g = a.g;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ short tap_kh[MAX_TAPS_TABLE];
  __shared__ short tap_kw[MAX_TAPS_TABLE];
  __shared__ int s_ntaps;

  const int tid = threadIdx.x;
  // ---- phase (conv-transpose form only): destination positions with (hd % sh, wd % sw) == (rh, rw)
  int rh = 0, rw = 0, step_h = 1, step_w = 1;
  if (g.transposed) {
    rh = blockIdx.z / g.sw;
    rw = blockIdx.z % g.sw;
    step_h = g.sh;
    step_w = g.sw;
  }
  const int Hp = (g.Hd - rh + step_h - 1) / step_h;
  const int Wp = (g.Wd - rw + step_w - 1) / step_w;
  const int64_t M = (int64_t)g.B * Hp * Wp;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * BN;

  int ntaps = g.KH * g.KW;
  if (g.transposed) {
    if (tid == 0) {
      // valid taps of a phase form an arithmetic progression with period s / gcd(d, s) along each axis
      auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
      auto first_valid = [](int r, int p, int d, int s, int period, int K) {
        for (int k = 0; k < period && k < K; ++k) {
          int t = r + p - k * d;
          if (((t % s) + s) % s == 0) return k;
        }
        return K;  // none
      };
      const int per_h = g.sh / gcd(g.dh % g.sh == 0 ? g.sh : g.dh % g.sh, g.sh);
      const int per_w = g.sw / gcd(g.dw % g.sw == 0 ? g.sw : g.dw % g.sw, g.sw);
      const int kh0 = first_valid(rh, g.ph, g.dh, g.sh, per_h, g.KH);
      const int kw0 = first_valid(rw, g.pw, g.dw, g.sw, per_w, g.KW);
      int c = 0;
      for (int kh = kh0; kh < g.KH; kh += per_h)
        for (int kw = kw0; kw < g.KW; kw += per_w) {
          tap_kh[c] = (short)kh;
          tap_kw[c] = (short)kw;
          ++c;
        }
      s_ntaps = c;
    }
    __syncthreads();
    ntaps = s_ntaps;
  }
  const int64_t Ktot = (int64_t)ntaps * g.Cs;

  // ---- per-thread A row (fixed over the K loop)
  const int a_row = tid >> 1;
  const int a_kpart = (tid & 1) * 8;
  const int64_t am = m0 + a_row;
  const bool a_row_ok = am < M;
  int ab = 0, ahd = 0, awd = 0;
  if (a_row_ok) {
    ab = (int)((unsigned)am / (unsigned)(Hp * Wp));
    int rem = (int)((unsigned)am - (unsigned)ab * (unsigned)(Hp * Wp));
    ahd = rh + (rem / Wp) * step_h;
    awd = rw + (rem % Wp) * step_w;
  }
  // ---- per-thread B coordinates
  int b_kk, b_nn;
  if (a.b_kmajor) { b_kk = tid & 15; b_nn = (tid >> 4) * TN; }
  else            { b_kk = tid >> 4; b_nn = (tid & 15) * TN; }

  float areg[8];
  float breg[TN];
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;
  const bool need_aux = xf_needs_aux(g.src_xf);

  auto decode_tap = [&](int t, int& kh, int& kw) {
    if (g.transposed) { kh = tap_kh[t]; kw = tap_kw[t]; }
    else { kh = t / g.KW; kw = t - kh * g.KW; }
  };

  auto load_tiles = [&](int64_t k0) {
    // A: 8 consecutive reduction indices of one output row
    const int64_t kb = k0 + a_kpart;
#pragma unroll
    for (int i = 0; i < 8; ++i) areg[i] = 0.f;
    if (a_row_ok && kb < Ktot) {
      if (a.vec_src) {
        int t = (int)(kb / g.Cs);
        int c = (int)(kb - (int64_t)t * g.Cs);
        int kh, kw, hs, ws;
        decode_tap(t, kh, kw);
        if (src_coord(g, ahd, awd, kh, kw, hs, ws)) {
          const int64_t off = (((int64_t)ab * g.Hs + hs) * g.Ws + ws);
          const float4* p = reinterpret_cast<const float4*>(a.src + off * g.ld_src + c);
          float4 v0 = __ldg(p), v1 = __ldg(p + 1);
          areg[0] = v0.x; areg[1] = v0.y; areg[2] = v0.z; areg[3] = v0.w;
          areg[4] = v1.x; areg[5] = v1.y; areg[6] = v1.z; areg[7] = v1.w;
          if (g.src_xf != MSMC_XF_NONE) {
            float aux[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (need_aux) {
              const float4* q = reinterpret_cast<const float4*>(a.src_aux + off * g.ld_saux + c);
              float4 u0 = __ldg(q), u1 = __ldg(q + 1);
              aux[0] = u0.x; aux[1] = u0.y; aux[2] = u0.z; aux[3] = u0.w;
              aux[4] = u1.x; aux[5] = u1.y; aux[6] = u1.z; aux[7] = u1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) areg[i] = apply_xf(g.src_xf, g.src_slope, areg[i], aux[i]);
          }
        }
      } else {
        int t = (int)(kb / g.Cs);
        int c = (int)(kb - (int64_t)t * g.Cs);
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          if (kb + i < Ktot) {
            int kh, kw, hs, ws;
            decode_tap(t, kh, kw);
            if (src_coord(g, ahd, awd, kh, kw, hs, ws)) {
              const int64_t off = (((int64_t)ab * g.Hs + hs) * g.Ws + ws);
              float v = __ldg(a.src + off * g.ld_src + c);
              float aux = need_aux ? __ldg(a.src_aux + off * g.ld_saux + c) : 0.f;
              areg[i] = apply_xf(g.src_xf, g.src_slope, v, aux);
            }
          }
          if (++c == g.Cs) { c = 0; ++t; }
        }
      }
    }
    // B: weights
    const int64_t k = k0 + b_kk;
#pragma unroll
    for (int j = 0; j < TN; ++j) breg[j] = 0.f;
    if (k < Ktot) {
      int t = (int)(k / g.Cs);
      int c = (int)(k - (int64_t)t * g.Cs);
      int kh, kw;
      decode_tap(t, kh, kw);
      const float* wp = a.w + kh * g.ws_kh + kw * g.ws_kw + c * g.ws_cs;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int n = n0 + b_nn + j;
        if (n < g.Cd) breg[j] = __ldg(wp + (int64_t)n * g.ws_cd);
      }
    }
  };

  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_kpart + i][a_row] = areg[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) Bs[b_kk][b_nn + j] = breg[j];
  };

  load_tiles(0);
  store_tiles();
  __syncthreads();
  for (int64_t k0 = 0; k0 < Ktot; k0 += BK) {
    const bool has_next = (k0 + BK) < Ktot;
    if (has_next) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[8], bv[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (has_next) {
      store_tiles();
      __syncthreads();
    }
  }

  // ---- epilogue: bias, result transform, residual
  const bool dneed_aux = xf_needs_aux(g.dst_xf);
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) break;
    const unsigned hw = (unsigned)(Hp * Wp);
    const int b = (int)((unsigned)m / hw);
    const int rem = (int)((unsigned)m - (unsigned)b * hw);
    const int hd = rh + (rem / Wp) * step_h;
    const int wd = rw + (rem % Wp) * step_w;
    const int64_t row = ((int64_t)b * g.Hd + hd) * g.Wd + wd;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.Cd) {
        float v = acc[i][j];
        if (a.bias) v += __ldg(a.bias + n);
        if (g.dst_xf != MSMC_XF_NONE) {
          float aux = dneed_aux ? __ldg(a.dst_aux + row * g.ld_daux + n) : 0.f;
          v = apply_xf(g.dst_xf, g.dst_slope, v, aux);
        }
        if (a.residual) v += __ldg(a.residual + row * g.ld_res + n);
        a.dst[row * g.ld_dst + n] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Few-channel convolutions (<= 16 input and <= 16 output channels: the first layers of the discriminators and
// their data gradients).  A 128 x 16 x 16 GEMM tile is mostly padding for them, so this kernel is direct: one
// thread per destination position holds all CDP output channels in registers, the whole weight sits in shared
// memory as [tap][cs][CDP] and is read with warp-broadcast 16-byte loads (one LDS.128 per 4 FMAs), the source
// patch comes through L1 (neighbouring threads read neighbouring pixels).  Forward and conv-transpose forms share
// src_coord(); taps that fall on padding or between stride phases are skipped.
// ------------------------------------------------------------------------------------------------
constexpr int DS_THREADS = 128;
constexpr int DS_MAX_W = 2304;   // (tap, cs, CDP) weight floats in shared memory

template <int CDP>
__global__ void __launch_bounds__(DS_THREADS) conv_direct_small_kernel(const ConvArgs a) {
  const msmc_conv_geom& g = a.g;
  __shared__ __align__(16) float s_w[DS_MAX_W];
  const int T = g.KH * g.KW;
  for (int e = threadIdx.x; e < T * g.Cs * CDP; e += DS_THREADS) {
    const int n = e % CDP;
    const int r = e / CDP;
    const int c = r % g.Cs, t = r / g.Cs;
    const int kh = t / g.KW, kw = t - kh * g.KW;
    s_w[e] = n < g.Cd ? __ldg(a.w + kh * g.ws_kh + kw * g.ws_kw + c * g.ws_cs + (int64_t)n * g.ws_cd) : 0.f;
  }
  __syncthreads();
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int64_t m = (int64_t)blockIdx.x * DS_THREADS + threadIdx.x;
  if (m >= M) return;
  const unsigned hw = (unsigned)(g.Hd * g.Wd);
  const int b = (int)((unsigned)m / hw);
  const int rem = (int)((unsigned)m - (unsigned)b * hw);
  const int hd = (int)((unsigned)rem / (unsigned)g.Wd), wd = rem - hd * g.Wd;
  const bool need_aux = xf_needs_aux(g.src_xf);

  float acc[CDP];
#pragma unroll
  for (int j = 0; j < CDP; ++j) acc[j] = 0.f;
  for (int kh = 0; kh < g.KH; ++kh)
    for (int kw = 0; kw < g.KW; ++kw) {
      int hs, ws;
      if (!src_coord(g, hd, wd, kh, kw, hs, ws)) continue;
      const int64_t off = ((int64_t)b * g.Hs + hs) * g.Ws + ws;
      const float* sp = a.src + off * g.ld_src;
      const float* ap = need_aux ? a.src_aux + off * g.ld_saux : nullptr;
      const float* wt = s_w + (kh * g.KW + kw) * g.Cs * CDP;
      auto mac = [&](float v, float aux, int c) {
        if (g.src_xf != MSMC_XF_NONE) v = apply_xf(g.src_xf, g.src_slope, v, aux);
        const float4* w4 = reinterpret_cast<const float4*>(wt + c * CDP);
#pragma unroll
        for (int j = 0; j < CDP / 4; ++j) {
          const float4 w = w4[j];
          acc[4 * j + 0] = fmaf(v, w.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
        }
      };
      if (a.vec_src) {          // Cs % 4 == 0 here, 16-byte aligned rows
        for (int c = 0; c < g.Cs; c += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(sp + c));
          float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
          if (need_aux) u = __ldg(reinterpret_cast<const float4*>(ap + c));
          mac(v.x, u.x, c); mac(v.y, u.y, c + 1); mac(v.z, u.z, c + 2); mac(v.w, u.w, c + 3);
        }
      } else {
        for (int c = 0; c < g.Cs; ++c) mac(__ldg(sp + c), need_aux ? __ldg(ap + c) : 0.f, c);
      }
    }
  const bool dneed_aux = xf_needs_aux(g.dst_xf);
#pragma unroll
  for (int n = 0; n < CDP; ++n) {
    if (n < g.Cd) {
      float v = acc[n];
      if (a.bias) v += __ldg(a.bias + n);
      if (g.dst_xf != MSMC_XF_NONE)
        v = apply_xf(g.dst_xf, g.dst_slope, v, dneed_aux ? __ldg(a.dst_aux + m * g.ld_daux + n) : 0.f);
      if (a.residual) v += __ldg(a.residual + m * g.ld_res + n);
      acc[n] = v;
    }
  }
  float* out = a.dst + m * g.ld_dst;
  if ((g.Cd & 3) == 0 && (g.ld_dst & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dst) & 15) == 0) {
#pragma unroll
    for (int n = 0; n < CDP; n += 4)
      if (n < g.Cd) *reinterpret_cast<float4*>(out + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
  } else {
#pragma unroll
    for (int n = 0; n < CDP; ++n)
      if (n < g.Cd) out[n] = acc[n];
  }
}

// ------------------------------------------------------------------------------------------------
// One source channel, many destination channels: the data gradients of the discriminators' score convolutions
// (C -> 1 forward, hence 1 -> C in conv-transpose form) and of the generator's conv_post.  The contraction is an outer
// product per tap (K = taps x 1), so a GEMM tile is all padding: here one thread owns 4 destination channels of one
// position, the weight [tap][Cd] sits in shared memory, the few source scalars come through L1 (the Cd/4 threads of a
// position read the same addresses) and the store is one coalesced 16-byte write per thread -- the kernel is bound
// by writing dst.  src_coord() gives forward and conv-transpose forms.
// ------------------------------------------------------------------------------------------------
constexpr int C1_THREADS = 256;
constexpr int C1_MAX_W = 8192;   // taps x Cd weight floats in shared memory

__global__ void __launch_bounds__(C1_THREADS) conv_c1_kernel(const ConvArgs a) {
  const msmc_conv_geom& g = a.g;
  __shared__ __align__(16) float s_w[C1_MAX_W];
  const int T = g.KH * g.KW;
  for (int e = threadIdx.x; e < T * g.Cd; e += C1_THREADS) {
    const int n = e % g.Cd, t = e / g.Cd;
    const int kh = t / g.KW, kw = t - kh * g.KW;
    s_w[e] = __ldg(a.w + kh * g.ws_kh + kw * g.ws_kw + (int64_t)n * g.ws_cd);
  }
  __syncthreads();
  const int cd4 = g.Cd >> 2;                       // <= 256
  const int P = C1_THREADS / cd4;                  // positions per CTA pass
  const int cg = threadIdx.x % cd4, pl = threadIdx.x / cd4;
  if (pl >= P) return;
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const unsigned hw = (unsigned)(g.Hd * g.Wd);
  const bool need_aux = xf_needs_aux(g.src_xf), dneed_aux = xf_needs_aux(g.dst_xf);
  const float4* w4 = reinterpret_cast<const float4*>(s_w);
  for (int64_t m = (int64_t)blockIdx.x * P + pl; m < M; m += (int64_t)gridDim.x * P) {
    const int b = (int)((unsigned)m / hw);
    const int rem = (int)((unsigned)m - (unsigned)b * hw);
    const int hd = (int)((unsigned)rem / (unsigned)g.Wd), wd = rem - hd * g.Wd;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw) {
        int hs, ws;
        if (!src_coord(g, hd, wd, kh, kw, hs, ws)) continue;
        const int64_t off = ((int64_t)b * g.Hs + hs) * g.Ws + ws;
        float v = __ldg(a.src + off * g.ld_src);
        if (g.src_xf != MSMC_XF_NONE)
          v = apply_xf(g.src_xf, g.src_slope, v, need_aux ? __ldg(a.src_aux + off * g.ld_saux) : 0.f);
        const float4 w = w4[(kh * g.KW + kw) * cd4 + cg];
        acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y);
        acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
      }
    const int n = cg * 4;
    if (a.bias) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + n));
      acc.x += bv.x; acc.y += bv.y; acc.z += bv.z; acc.w += bv.w;
    }
    if (g.dst_xf != MSMC_XF_NONE) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dneed_aux) av = __ldg(reinterpret_cast<const float4*>(a.dst_aux + m * g.ld_daux + n));
      acc.x = apply_xf(g.dst_xf, g.dst_slope, acc.x, av.x); acc.y = apply_xf(g.dst_xf, g.dst_slope, acc.y, av.y);
      acc.z = apply_xf(g.dst_xf, g.dst_slope, acc.z, av.z); acc.w = apply_xf(g.dst_xf, g.dst_slope, acc.w, av.w);
    }
    if (a.residual) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(a.residual + m * g.ld_res + n));
      acc.x += rv.x; acc.y += rv.y; acc.z += rv.z; acc.w += rv.w;
    }
    *reinterpret_cast<float4*>(a.dst + m * g.ld_dst + n) = acc;
  }
}

bool direct_small_eligible(const msmc_conv_geom& g) {
  if (g.Cs > 16 || g.Cd > 16) return false;
  const int cdp = g.Cd <= 4 ? 4 : (g.Cd <= 8 ? 8 : 16);
  return (int64_t)g.KH * g.KW * g.Cs * cdp <= DS_MAX_W;
}

// ------------------------------------------------------------------------------------------------
// weight gradient:  dW[(tap,cs), cd] = sum_m  xf(src)[m,(tap,cs)] * xf(gout)[m, cd]
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* gout;
  const float* gout_aux;
  float* partial;         // [splits][Ktot*Cd + Cd]
  int64_t rows_per_split; // multiple of BK
  int vec_src;
  int want_bias;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS) conv_wgrad_kernel(const WgradArgs a) {
  constexpr int TN = BN / 16;
  const msmc_conv_geom& g = a.g;
  __shared__ __align__(16) float As[BK][BM + 4];   // [row m within chunk][reduction-index tile]
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ float bias_red[BK][BN + 4];

  const int tid = threadIdx.x;
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int64_t kt0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t mbeg = (int64_t)blockIdx.z * a.rows_per_split;
  const int64_t mend = min(M, mbeg + a.rows_per_split);

  const int l_mm = tid >> 4;              // row within the chunk (0..15)
  const int l_kpart = (tid & 15) * 8;     // 8 consecutive reduction indices
  const int l_nn = (tid & 15) * TN;
  const int ty = tid >> 4, tx = tid & 15;
  const bool need_aux = xf_needs_aux(g.src_xf);
  const bool gneed_aux = xf_needs_aux(g.dst_xf);
  const bool do_bias = a.want_bias && blockIdx.x == 0;

  float areg[8], breg[TN], bsum[TN];
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < TN; ++j) bsum[j] = 0.f;

  // the (tap, channel) of this thread's 8 reduction indices never changes over the row loop
  const int64_t kb = kt0 + l_kpart;

  auto load_tiles = [&](int64_t mb) {
    const int64_t m = mb + l_mm;
#pragma unroll
    for (int i = 0; i < 8; ++i) areg[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) breg[j] = 0.f;
    if (m < mend) {
      // 32-bit unsigned division (host guarantees M < 2^31): ~10x cheaper than the 64-bit one
      const unsigned hw = (unsigned)(g.Hd * g.Wd);
      const int b = (int)((unsigned)m / hw);
      const int rem = (int)((unsigned)m - (unsigned)b * hw);
      const int hd = (int)((unsigned)rem / (unsigned)g.Wd), wd = rem - hd * g.Wd;
      if (kb < Ktot) {
        int t = (int)(kb / g.Cs);
        int c = (int)(kb - (int64_t)t * g.Cs);
        if (a.vec_src) {
          int kh = t / g.KW, kw = t - kh * g.KW, hs, ws;
          if (src_coord(g, hd, wd, kh, kw, hs, ws)) {
            const int64_t off = (((int64_t)b * g.Hs + hs) * g.Ws + ws);
            const float4* p = reinterpret_cast<const float4*>(a.src + off * g.ld_src + c);
            float4 v0 = __ldg(p), v1 = __ldg(p + 1);
            areg[0] = v0.x; areg[1] = v0.y; areg[2] = v0.z; areg[3] = v0.w;
            areg[4] = v1.x; areg[5] = v1.y; areg[6] = v1.z; areg[7] = v1.w;
            if (g.src_xf != MSMC_XF_NONE) {
              float aux[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              if (need_aux) {
                const float4* q = reinterpret_cast<const float4*>(a.src_aux + off * g.ld_saux + c);
                float4 u0 = __ldg(q), u1 = __ldg(q + 1);
                aux[0] = u0.x; aux[1] = u0.y; aux[2] = u0.z; aux[3] = u0.w;
                aux[4] = u1.x; aux[5] = u1.y; aux[6] = u1.z; aux[7] = u1.w;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) areg[i] = apply_xf(g.src_xf, g.src_slope, areg[i], aux[i]);
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            if (kb + i < Ktot) {
              int kh = t / g.KW, kw = t - kh * g.KW, hs, ws;
              if (src_coord(g, hd, wd, kh, kw, hs, ws)) {
                const int64_t off = (((int64_t)b * g.Hs + hs) * g.Ws + ws);
                float v = __ldg(a.src + off * g.ld_src + c);
                float aux = need_aux ? __ldg(a.src_aux + off * g.ld_saux + c) : 0.f;
                areg[i] = apply_xf(g.src_xf, g.src_slope, v, aux);
              }
            }
            if (++c == g.Cs) { c = 0; ++t; }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + l_nn + j;
        if (n < g.Cd) {
          float v = __ldg(a.gout + m * g.ld_dst + n);
          if (g.dst_xf != MSMC_XF_NONE) {
            float aux = gneed_aux ? __ldg(a.gout_aux + m * g.ld_daux + n) : 0.f;
            v = apply_xf(g.dst_xf, g.dst_slope, v, aux);
          }
          breg[j] = v;
        }
      }
    }
  };
  auto store_tiles = [&]() {
    *reinterpret_cast<float4*>(&As[l_mm][l_kpart]) = make_float4(areg[0], areg[1], areg[2], areg[3]);
    *reinterpret_cast<float4*>(&As[l_mm][l_kpart + 4]) = make_float4(areg[4], areg[5], areg[6], areg[7]);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      Bs[l_mm][l_nn + j] = breg[j];
      bsum[j] += breg[j];
    }
  };

  if (mbeg < mend) {
    load_tiles(mbeg);
    store_tiles();
  }
  __syncthreads();
  for (int64_t mb = mbeg; mb < mend; mb += BK) {
    const bool has_next = (mb + BK) < mend;
    if (has_next) load_tiles(mb + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[8], bv[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (has_next) {
      store_tiles();
      __syncthreads();
    }
  }

  float* part = a.partial + (int64_t)blockIdx.z * (Ktot * g.Cd + g.Cd);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t k = kt0 + ty * 8 + i;
    if (k < Ktot) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < g.Cd) part[k * g.Cd + n] = acc[i][j];
      }
    }
  }
  if (do_bias) {
    // fixed-order reduction over the 16 row-lanes
#pragma unroll
    for (int j = 0; j < TN; ++j) bias_red[l_mm][l_nn + j] = bsum[j];
    __syncthreads();
    if (tid < BN) {
      float s = 0.f;
      for (int r = 0; r < BK; ++r) s += bias_red[r][tid];
      const int n = n0 + tid;
      if (n < g.Cd) part[Ktot * g.Cd + n] = s;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// weight gradient of the few-channel convolutions (first layers of the period / resolution discriminators:
// 1-8 input channels, 4-16 output channels, 10^5-10^6 positions).  The GEMM tile kernels above waste >90 % of a
// 128 x 16 tile on these; here the whole dW (plus the bias row) is one accumulator set per CTA:
//   stage 128 positions transposed in shared memory  s_x[k][p] (k = (tap, cs), row Ktot = ones -> bias),  s_g[cd][p]
//   thread (group, o) accumulates  dW[o] += sum_p s_x[k(o)][p] * s_g[cd(o)][p]  over its group's positions
// Partials per position-slice use the same [Ktot*Cd + Cd] workspace rows and second pass as the other kernels.
// ------------------------------------------------------------------------------------------------
constexpr int SW_TP = 128;               // positions per staged tile
constexpr int SW_PITCH = SW_TP + 4;      // +4 floats: 16-byte aligned rows, rows land 4 banks apart
constexpr int SW_THREADS = 256;
constexpr int SW_MAX_NO = 5;             // outputs per thread when (Ktot+1)*Cd > 256

struct SmallWgradArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* gout;
  const float* gout_aux;
  float* partial;          // [splits][Ktot*Cd + Cd]
  int64_t rows_per_split;  // multiple of SW_TP
  int groups;              // position groups per tile (power of two, 1..8)
  int outs_per_thread;     // 1..SW_MAX_NO
};

__global__ void __launch_bounds__(SW_THREADS) conv_wgrad_small_kernel(const SmallWgradArgs a) {
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(16) float sw_smem[];
  const int Ktot = g.KH * g.KW * g.Cs;
  const int O = (Ktot + 1) * g.Cd;
  float* s_x = sw_smem;                               // [(Ktot + 1)][SW_PITCH]
  float* s_g = sw_smem + (Ktot + 1) * SW_PITCH;       // [Cd][SW_PITCH]
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int64_t mbeg = (int64_t)blockIdx.x * a.rows_per_split;
  const int64_t mend = min(M, mbeg + a.rows_per_split);
  const bool need_aux = xf_needs_aux(g.src_xf);
  const bool gneed_aux = xf_needs_aux(g.dst_xf);

  const int tg = SW_THREADS / a.groups;               // threads per group
  const int grp = tid / tg, to = tid - grp * tg;
  const int span = SW_TP / a.groups;                  // positions per group and tile (multiple of 4)
  int xrow[SW_MAX_NO], grow[SW_MAX_NO];
  float acc[SW_MAX_NO];
#pragma unroll
  for (int j = 0; j < SW_MAX_NO; ++j) {
    const int o = to + j * tg;
    const bool ok = j < a.outs_per_thread && o < O;
    const int k = ok ? o / g.Cd : 0;
    xrow[j] = ok ? k * SW_PITCH : -1;
    grow[j] = ok ? (o - k * g.Cd) * SW_PITCH : 0;
    acc[j] = 0.f;
  }

  const int lp = tid & (SW_TP - 1);                   // position this thread gathers
  const int lk = tid / SW_TP;                         // 0 / 1: even / odd reduction rows
  const unsigned hw = (unsigned)(g.Hd * g.Wd);
  for (int64_t m0 = mbeg; m0 < mend; m0 += SW_TP) {
    // ---- stage the source patch: one position per thread, every other (tap, channel) row ----
    {
      const int64_t m = m0 + lp;
      const bool m_ok = m < mend;
      int b = 0, hd = 0, wd = 0;
      if (m_ok) {
        b = (int)((unsigned)m / hw);
        const int rem = (int)((unsigned)m - (unsigned)b * hw);
        hd = (int)((unsigned)rem / (unsigned)g.Wd);
        wd = rem - hd * g.Wd;
      }
      int t = 0, c = lk;                              // k = t * Cs + c, advanced by 2 per iteration
      while (c >= g.Cs) { c -= g.Cs; ++t; }
      for (int k = lk; k < Ktot; k += 2) {
        float v = 0.f;
        if (m_ok) {
          const int kh = t / g.KW, kw = t - kh * g.KW;
          int hs, ws;
          if (src_coord(g, hd, wd, kh, kw, hs, ws)) {
            const int64_t off = ((int64_t)b * g.Hs + hs) * g.Ws + ws;
            v = __ldg(a.src + off * g.ld_src + c);
            if (g.src_xf != MSMC_XF_NONE)
              v = apply_xf(g.src_xf, g.src_slope, v, need_aux ? __ldg(a.src_aux + off * g.ld_saux + c) : 0.f);
          }
        }
        s_x[k * SW_PITCH + lp] = v;
        c += 2;
        while (c >= g.Cs) { c -= g.Cs; ++t; }
      }
      if (lk == 0) s_x[Ktot * SW_PITCH + lp] = m_ok ? 1.f : 0.f;
    }
    // ---- stage the output gradient transposed ----
    for (int e = tid; e < SW_TP * g.Cd; e += SW_THREADS) {
      const int p = e / g.Cd, cd = e - p * g.Cd;
      const int64_t m = m0 + p;
      float v = 0.f;
      if (m < mend) {
        v = __ldg(a.gout + m * g.ld_dst + cd);
        if (g.dst_xf != MSMC_XF_NONE)
          v = apply_xf(g.dst_xf, g.dst_slope, v, gneed_aux ? __ldg(a.gout_aux + m * g.ld_daux + cd) : 0.f);
      }
      s_g[cd * SW_PITCH + p] = v;
    }
    __syncthreads();
    const int p0 = grp * span;
#pragma unroll
    for (int j = 0; j < SW_MAX_NO; ++j) {
      if (xrow[j] >= 0) {
        const float* xr = s_x + xrow[j] + p0;
        const float* gr = s_g + grow[j] + p0;
        float s = acc[j];
        for (int p = 0; p < span; p += 4) {
          const float4 x4 = *reinterpret_cast<const float4*>(xr + p);
          const float4 g4 = *reinterpret_cast<const float4*>(gr + p);
          s = fmaf(x4.x, g4.x, s); s = fmaf(x4.y, g4.y, s); s = fmaf(x4.z, g4.z, s); s = fmaf(x4.w, g4.w, s);
        }
        acc[j] = s;
      }
    }
    __syncthreads();
  }

  float* part = a.partial + (int64_t)blockIdx.x * O;
  if (a.groups == 1) {
#pragma unroll
    for (int j = 0; j < SW_MAX_NO; ++j)
      if (xrow[j] >= 0) part[to + j * tg] = acc[j];
  } else {
    // fixed-order sum over the position groups (outs_per_thread == 1 here)
    float* s_red = sw_smem;                            // [groups][tg], groups * tg = 256 floats
    s_red[grp * tg + to] = acc[0];
    __syncthreads();
    if (grp == 0 && to < O) {
      float s = 0.f;
      for (int q = 0; q < a.groups; ++q) s += s_red[q * tg + to];
      part[to] = s;
    }
  }
}

bool small_wgrad_plan(const msmc_conv_geom& g, int* groups, int* outs_per_thread) {
  if (g.transposed) return false;
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  if (Ktot > 80 || g.Cs > 8 || g.Cd > 16) return false;
  const int O = (int)(Ktot + 1) * g.Cd;
  if (O > 2 * SW_THREADS) return false;          // beyond two outputs per thread the 128 x 16 tile kernel wins
  int ng = 1;
  while (ng < 8 && SW_THREADS / (ng * 2) >= O) ng *= 2;
  *groups = ng;
  *outs_per_thread = ng > 1 ? 1 : ceil_div(O, SW_THREADS);
  return true;
}
int small_wgrad_splits(const msmc_conv_geom& g) {
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(M, 2 * SW_TP), (int64_t)num_sms() * 4));
}

// 64 outputs per CTA x 4 slice groups: group y sums slices y, y + 4, ... (four loads in flight), the groups are
// combined in fixed order through shared memory -- deterministic, and a quarter of the serial walk over the slices
// that bounds these small launches (12.5 us average, 204 per step, on the weight-gradient join of every conv backward)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const msmc_conv_geom g, const float* __restrict__ partial, int splits,
                    float* __restrict__ dw, float* __restrict__ dbias) {
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  const int64_t per = Ktot * g.Cd + g.Cd;
  const int64_t total = dbias ? per : Ktot * g.Cd;
  __shared__ float red[4][64];
  const int ex = threadIdx.x & 63, gy = threadIdx.x >> 6;
  for (int64_t e0 = (int64_t)blockIdx.x * 64; e0 < total; e0 += (int64_t)gridDim.x * 64) {
    const int64_t e = e0 + ex;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (e < total) {
      int z = gy;
      for (; z + 12 < splits; z += 16) {
        s0 += partial[(int64_t)z * per + e];
        s1 += partial[(int64_t)(z + 4) * per + e];
        s2 += partial[(int64_t)(z + 8) * per + e];
        s3 += partial[(int64_t)(z + 12) * per + e];
      }
      for (; z < splits; z += 4) s0 += partial[(int64_t)z * per + e];
    }
    red[gy][ex] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (gy == 0 && e < total) {
      const float s = ((red[0][ex] + red[1][ex]) + red[2][ex]) + red[3][ex];
      if (e < Ktot * g.Cd) {
        const int64_t k = e / g.Cd;
        const int n = (int)(e - k * g.Cd);
        const int t = (int)(k / g.Cs);
        const int c = (int)(k - (int64_t)t * g.Cs);
        const int kh = t / g.KW, kw = t - kh * g.KW;
        dw[kh * g.ws_kh + kw * g.ws_kw + c * g.ws_cs + (int64_t)n * g.ws_cd] = s;
      } else {
        dbias[e - Ktot * g.Cd] = s;
      }
    }
    __syncthreads();
  }
}

// weight_norm forward: w = g * v / ||v||_row, written through arbitrary output strides (GEMM layout)
__global__ void weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ gpar,
                                       float* __restrict__ w, float* __restrict__ inv_norm,
                                       int O, int I, int J, int64_t so, int64_t si, int64_t sj) {
  const int o = blockIdx.x;
  const int n = I * J;
  const float* vr = v + (int64_t)o * n;
  __shared__ float red[32];
  __shared__ float s_scale;
  float scale = 1.f;
  if (gpar) {
    float ss = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) { float x = vr[e]; ss = fmaf(x, x, ss); }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
      float inv = 1.f / sqrtf(t);
      if (inv_norm) inv_norm[o] = inv;
      s_scale = gpar[o] * inv;
    }
    __syncthreads();
    scale = s_scale;
  }
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e / J, j = e - i * J;
    w[o * so + i * si + j * sj] = vr[e] * scale;
  }
}

// weight_norm backward: dg = (dw . v) / ||v|| ;  dv = g/||v|| * (dw - v * (dw . v)/||v||^2)
__global__ void weight_norm_bwd_kernel(const float* __restrict__ dw, int64_t so, int64_t si, int64_t sj,
                                       const float* __restrict__ v, const float* __restrict__ gpar,
                                       const float* __restrict__ inv_norm, float* __restrict__ dv,
                                       float* __restrict__ dg, int O, int I, int J) {
  const int o = blockIdx.x;
  const int n = I * J;
  const float* vr = v + (int64_t)o * n;
  float* dvr = dv + (int64_t)o * n;
  if (!gpar) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int i = e / J, j = e - i * J;
      dvr[e] = dw[o * so + i * si + j * sj];
    }
    return;
  }
  __shared__ float red[32];
  __shared__ float s_dot;
  float dot = 0.f;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e / J, j = e - i * J;
    dot = fmaf(dw[o * so + i * si + j * sj], vr[e], dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    s_dot = t;
    dg[o] = t * inv_norm[o];
  }
  __syncthreads();
  const float inv = inv_norm[o];
  const float gs = gpar[o] * inv;
  const float coef = s_dot * inv * inv;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e / J, j = e - i * J;
    dvr[e] = gs * (dw[o * so + i * si + j * sj] - vr[e] * coef);
  }
}

// backward of ReflectionPad: fold a (B, H+2p, W+2q, C) gradient onto (B, H, W, C)
// VEC channels per thread (4 when C is a multiple of 4: one 16-byte load per contributing pixel), 32-bit index
// arithmetic (the scalar form spent ~100 instructions of 64-bit div / mod per element and ran at 0.6 TB/s)
template <int VEC>
__global__ void reflect_fold_kernel(const float* __restrict__ gp, float* __restrict__ gx, int B, int H, int W,
                                    int C, int ph, int pw) {
  const int CV = C / VEC;
  const uint32_t total = (uint32_t)B * H * W * CV;      // (checked < 2^31 by the launcher)
  const int Hp = H + 2 * ph, Wp = W + 2 * pw;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const uint32_t cv = e % (uint32_t)CV;
    uint32_t r = e / (uint32_t)CV;
    const int w = (int)(r % (uint32_t)W); r /= (uint32_t)W;
    const int h = (int)(r % (uint32_t)H);
    const int b = (int)(r / (uint32_t)H);
    // padded coordinates that reflect onto (h, w): itself, and mirror images near each border
    int hc[3], wc[3], nh = 0, nw = 0;
    hc[nh++] = h + ph;
    if (h >= 1 && h <= ph) hc[nh++] = ph - h;
    if (h <= H - 2 && h >= H - 1 - ph) hc[nh++] = ph + 2 * (H - 1) - h;
    wc[nw++] = w + pw;
    if (w >= 1 && w <= pw) wc[nw++] = pw - w;
    if (w <= W - 2 && w >= W - 1 - pw) wc[nw++] = pw + 2 * (W - 1) - w;
    float s[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[v] = 0.f;
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nw; ++j) {
        const float* src = gp + (((int64_t)b * Hp + hc[i]) * Wp + wc[j]) * C + cv * VEC;
        if (VEC == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src));
          s[0] += t.x; s[1 % VEC] += t.y; s[2 % VEC] += t.z; s[3 % VEC] += t.w;
        } else {
          s[0] += __ldg(src);
        }
      }
    float* dst = gx + (int64_t)e * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1 % VEC], s[2 % VEC], s[3 % VEC]);
    else dst[0] = s[0];
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int pick_bn(int cd) { return cd <= 16 ? 16 : (cd <= 32 ? 32 : (cd <= 64 ? 64 : 128)); }

}  // namespace
}  // namespace msmc

namespace msmc {
int launch_wgrad_reduce(const msmc_conv_geom& g, const float* workspace, int splits, float* dw, float* dbias,
                        void* stream) {
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  const int64_t total = Ktot * g.Cd + g.Cd;
  // (a 4-wide variant -- one 16-byte load per slice, a quarter of the threads -- measured 2x slower per launch:
  //  these launches are latency-bound on the serial walk over the slices and live on thread count)
  int blocks = (int)std::min<int64_t>(ceil_div64(total, 64), (int64_t)num_sms() * 8);
  wgrad_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, workspace, splits, dw, dbias);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
}  // namespace msmc

using namespace msmc;

extern "C" int msmc_conv_forward(const msmc_conv_geom* gp, const float* src, const float* src_aux,
                                 const float* w, const float* bias, const float* residual,
                                 const float* dst_aux, float* dst, void* stream) {
  MSMC_REQUIRE(gp && src && w && dst);
  const msmc_conv_geom& g = *gp;
  MSMC_REQUIRE(g.B > 0 && g.Cs > 0 && g.Cd > 0 && g.KH > 0 && g.KW > 0 && g.sh > 0 && g.sw > 0);
  MSMC_REQUIRE(g.Hs > 0 && g.Ws > 0 && g.Hd > 0 && g.Wd > 0);
  MSMC_REQUIRE((int64_t)g.B * g.Hd * g.Wd < ((int64_t)1 << 31) && (int64_t)g.B * g.Hs * g.Ws < ((int64_t)1 << 31));
  MSMC_REQUIRE(!(g.transposed && g.pad_reflect));
  MSMC_REQUIRE(!xf_needs_aux(g.src_xf) || src_aux);
  MSMC_REQUIRE(!xf_needs_aux(g.dst_xf) || dst_aux);
  if (g.transposed) MSMC_REQUIRE((int64_t)g.KH * g.KW <= MAX_TAPS_TABLE);
  if (g.pad_reflect) MSMC_REQUIRE((g.ph == 0 || g.ph < g.Hs) && (g.pw == 0 || g.pw < g.Ws));
  ConvArgs a;
  a.g = g; a.src = src; a.src_aux = src_aux; a.w = w; a.bias = bias; a.residual = residual;
  a.dst_aux = dst_aux; a.dst = dst;
  a.vec_src = (g.Cs % 8 == 0) && (g.ld_src % 4 == 0) && aligned16(src) &&
              (!xf_needs_aux(g.src_xf) || ((g.ld_saux % 4 == 0) && aligned16(src_aux)));
  a.b_kmajor = (g.ws_cd != 1 && g.ws_cs == 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (g.Cs == 1 && g.Cd >= 32 && g.Cd <= 1024 && (g.Cd & 3) == 0 && (int64_t)g.KH * g.KW * g.Cd <= C1_MAX_W &&
      (g.ld_dst & 3) == 0 && aligned16(dst) && (!bias || aligned16(bias)) &&
      (!residual || ((g.ld_res & 3) == 0 && aligned16(residual))) &&
      (!xf_needs_aux(g.dst_xf) || ((g.ld_daux & 3) == 0 && aligned16(dst_aux)))) {
    const int64_t Mall = (int64_t)g.B * g.Hd * g.Wd;
    const int P = C1_THREADS / (g.Cd >> 2);
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(Mall, P), (int64_t)num_sms() * 8);
    conv_c1_kernel<<<blocks, C1_THREADS, 0, st>>>(a);
    MSMC_CHECK_LAUNCH();
    return MSMC_OK;
  }
  if (direct_small_eligible(g)) {
    a.vec_src = (g.Cs % 4 == 0) && (g.ld_src % 4 == 0) && aligned16(src) &&
                (!xf_needs_aux(g.src_xf) || ((g.ld_saux % 4 == 0) && aligned16(src_aux)));
    const int64_t Mall = (int64_t)g.B * g.Hd * g.Wd;
    const unsigned blocks = (unsigned)ceil_div64(Mall, DS_THREADS);
    if (g.Cd <= 4) conv_direct_small_kernel<4><<<blocks, DS_THREADS, 0, st>>>(a);
    else if (g.Cd <= 8) conv_direct_small_kernel<8><<<blocks, DS_THREADS, 0, st>>>(a);
    else conv_direct_small_kernel<16><<<blocks, DS_THREADS, 0, st>>>(a);
    MSMC_CHECK_LAUNCH();
    return MSMC_OK;
  }
  const int phases = g.transposed ? g.sh * g.sw : 1;
  int64_t maxM;
  if (g.transposed) maxM = (int64_t)g.B * ceil_div(g.Hd, g.sh) * ceil_div(g.Wd, g.sw);
  else maxM = (int64_t)g.B * g.Hd * g.Wd;
  const int bn = pick_bn(g.Cd);
  dim3 grid((unsigned)ceil_div64(maxM, BM), (unsigned)ceil_div(g.Cd, bn), (unsigned)phases);
  switch (bn) {
    case 16: conv_gemm_kernel<16><<<grid, NTHREADS, 0, st>>>(a); break;
    case 32: conv_gemm_kernel<32><<<grid, NTHREADS, 0, st>>>(a); break;
    case 64: conv_gemm_kernel<64><<<grid, NTHREADS, 0, st>>>(a); break;
    default: conv_gemm_kernel<128><<<grid, NTHREADS, 0, st>>>(a); break;
  }
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

namespace {
int wgrad_splits(const msmc_conv_geom& g) {
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int bn = msmc::pick_bn(g.Cd);
  const int64_t tiles = ceil_div64(Ktot, BM) * ceil_div(g.Cd, bn);
  const int64_t target = (int64_t)msmc::num_sms() * 3;
  int64_t s = ceil_div64(target, tiles);
  const int64_t max_by_rows = ceil_div64(M, 4 * BK);
  if (s > max_by_rows) s = max_by_rows;
  // bound the partial-sum workspace to 96 MiB
  const int64_t per = (Ktot * g.Cd + g.Cd) * (int64_t)sizeof(float);
  const int64_t cap = ((int64_t)96 << 20) / per;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return (int)s;
}
}  // namespace

extern "C" int64_t msmc_conv_wgrad_workspace(const msmc_conv_geom* gp) {
  if (!gp) return -1;
  const msmc_conv_geom& g = *gp;
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  int ng, no;
  const int splits = msmc::small_wgrad_plan(g, &ng, &no) ? msmc::small_wgrad_splits(g) : wgrad_splits(g);
  return (int64_t)splits * (Ktot * g.Cd + g.Cd) * (int64_t)sizeof(float);
}

extern "C" int msmc_conv_wgrad(const msmc_conv_geom* gp, const float* src, const float* src_aux,
                               const float* gout, const float* gout_aux, float* dw, float* dbias,
                               float* workspace, int64_t workspace_bytes, void* stream) {
  MSMC_REQUIRE(gp && src && gout && dw && workspace);
  const msmc_conv_geom& g = *gp;
  MSMC_REQUIRE(!g.transposed);
  MSMC_REQUIRE((int64_t)g.B * g.Hd * g.Wd < ((int64_t)1 << 31) && (int64_t)g.B * g.Hs * g.Ws < ((int64_t)1 << 31));
  MSMC_REQUIRE(!xf_needs_aux(g.src_xf) || src_aux);
  MSMC_REQUIRE(!xf_needs_aux(g.dst_xf) || gout_aux);
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  cudaStream_t st = (cudaStream_t)stream;
  {
    SmallWgradArgs sa;
    if (small_wgrad_plan(g, &sa.groups, &sa.outs_per_thread)) {
      const int ssplits = small_wgrad_splits(g);
      MSMC_REQUIRE(workspace_bytes >= (int64_t)ssplits * (Ktot * g.Cd + g.Cd) * (int64_t)sizeof(float));
      sa.g = g; sa.src = src; sa.src_aux = src_aux; sa.gout = gout; sa.gout_aux = gout_aux; sa.partial = workspace;
      sa.rows_per_split = ceil_div64(ceil_div64(M, ssplits), SW_TP) * SW_TP;
      const int eff = (int)ceil_div64(M, sa.rows_per_split);
      const size_t smem = (size_t)std::max<int64_t>((Ktot + 1 + g.Cd) * SW_PITCH, SW_THREADS) * sizeof(float);
      static bool attr_done = false;
      if (!attr_done) {
        cudaFuncSetAttribute(conv_wgrad_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_done = true;
      }
      conv_wgrad_small_kernel<<<eff, SW_THREADS, smem, st>>>(sa);
      MSMC_CHECK_LAUNCH();
      // the bias row is always produced; the second pass ignores it when dbias is null
      return launch_wgrad_reduce(g, workspace, eff, dw, dbias, stream);
    }
  }
  const int splits = wgrad_splits(g);
  MSMC_REQUIRE(workspace_bytes >= (int64_t)splits * (Ktot * g.Cd + g.Cd) * (int64_t)sizeof(float));
  WgradArgs a;
  a.g = g; a.src = src; a.src_aux = src_aux; a.gout = gout; a.gout_aux = gout_aux;
  a.partial = workspace;
  a.rows_per_split = ceil_div64(ceil_div64(M, splits), BK) * BK;
  a.vec_src = (g.Cs % 8 == 0) && (g.ld_src % 4 == 0) && aligned16(src) &&
              (!xf_needs_aux(g.src_xf) || ((g.ld_saux % 4 == 0) && aligned16(src_aux)));
  a.want_bias = dbias != nullptr;
  const int bn = pick_bn(g.Cd);
  dim3 grid((unsigned)ceil_div64(Ktot, BM), (unsigned)ceil_div(g.Cd, bn), (unsigned)splits);
  switch (bn) {
    case 16: conv_wgrad_kernel<16><<<grid, NTHREADS, 0, st>>>(a); break;
    case 32: conv_wgrad_kernel<32><<<grid, NTHREADS, 0, st>>>(a); break;
    case 64: conv_wgrad_kernel<64><<<grid, NTHREADS, 0, st>>>(a); break;
    default: conv_wgrad_kernel<128><<<grid, NTHREADS, 0, st>>>(a); break;
  }
  MSMC_CHECK_LAUNCH();
  return launch_wgrad_reduce(g, workspace, splits, dw, dbias, stream);
}

extern "C" int msmc_weight_norm_fwd(const float* v, const float* g, float* w, float* inv_norm, int32_t O,
                                    int32_t I, int32_t J, int64_t so, int64_t si, int64_t sj, void* stream) {
  MSMC_REQUIRE(v && w && O > 0 && I > 0 && J > 0);
  weight_norm_fwd_kernel<<<O, 256, 0, (cudaStream_t)stream>>>(v, g, w, inv_norm, O, I, J, so, si, sj);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Multi-tensor weight preparation: ONE launch re-parametrises every conv / linear weight of a sub-network
// (weight_norm g * v / ||v|| fused with the re-layout to the GEMM layout), instead of one 3-10 us launch per layer
// (268 + 430 launches per train step in round 1, with msmc_weight_image_multi below).  Job table (device, int64
// words, 11 per job): v, g (0 = plain re-layout), w, inv_norm (0 = not wanted), so, si, sj, O, I, J, row0 where
// row0 = sum of O over the preceding jobs; CTA b serves global row b (binary search over row0).
// ------------------------------------------------------------------------------------------------------------
namespace msmc {
namespace {
constexpr int WN_JOB_WORDS = 11;
__global__ void __launch_bounds__(256) weight_norm_fwd_multi_kernel(const long long* __restrict__ jobs, int n_jobs) {
  int lo = 0, hi = n_jobs - 1;
  const long long row = blockIdx.x;
  while (lo < hi) {                       // last job with row0 <= row
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[(size_t)mid * WN_JOB_WORDS + 10] <= row) lo = mid; else hi = mid - 1;
  }
  const long long* jb = jobs + (size_t)lo * WN_JOB_WORDS;
  const float* v = reinterpret_cast<const float*>(jb[0]);
  const float* gpar = reinterpret_cast<const float*>(jb[1]);
  float* w = reinterpret_cast<float*>(jb[2]);
  float* inv_norm = reinterpret_cast<float*>(jb[3]);
  const long long so = jb[4], si = jb[5], sj = jb[6];
  const int I = (int)jb[8], J = (int)jb[9];
  const int o = (int)(row - jb[10]);
  const int n = I * J;
  const float* vr = v + (int64_t)o * n;
  __shared__ float red[32];
  __shared__ float s_scale;
  float scale = 1.f;
  if (gpar) {
    // identical arithmetic and order to weight_norm_fwd_kernel (same block size): bit-identical results
    float ss = 0.f;
    for (int e = threadIdx.x; e < n; e += blockDim.x) { float x = vr[e]; ss = fmaf(x, x, ss); }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
      float inv = 1.f / sqrtf(t);
      if (inv_norm) inv_norm[o] = inv;
      s_scale = gpar[o] * inv;
    }
    __syncthreads();
    scale = s_scale;
  }
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e / J, j = e - i * J;
    w[o * so + i * si + j * sj] = vr[e] * scale;
  }
}
}  // namespace
}  // namespace msmc

extern "C" int msmc_weight_norm_fwd_multi(const int64_t* jobs, int32_t n_jobs, int64_t total_rows, void* stream) {
  MSMC_REQUIRE(jobs && n_jobs > 0 && total_rows > 0 && total_rows < ((int64_t)1 << 31));
  weight_norm_fwd_multi_kernel<<<(unsigned)total_rows, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(jobs), n_jobs);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_weight_norm_bwd(const float* dw, int64_t so, int64_t si, int64_t sj, const float* v,
                                    const float* g, const float* inv_norm, float* dv, float* dg, int32_t O,
                                    int32_t I, int32_t J, void* stream) {
  MSMC_REQUIRE(dw && v && dv && O > 0 && I > 0 && J > 0);
  MSMC_REQUIRE(!g || (inv_norm && dg));
  weight_norm_bwd_kernel<<<O, 256, 0, (cudaStream_t)stream>>>(dw, so, si, sj, v, g, inv_norm, dv, dg, O, I, J);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_reflect_pad_fold(const float* gpad, float* gx, int32_t B, int32_t H, int32_t W, int32_t C,
                                     int32_t ph, int32_t pw, void* stream) {
  MSMC_REQUIRE(gpad && gx && B > 0 && H > 0 && W > 0 && C > 0 && ph >= 0 && pw >= 0);
  MSMC_REQUIRE((ph < H || ph == 0) && (pw < W || pw == 0));
  const bool vec4 = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(gpad) | reinterpret_cast<uintptr_t>(gx)) & 15) == 0;
  const int64_t total = (int64_t)B * H * W * (vec4 ? C / 4 : C);
  MSMC_REQUIRE(total < (int64_t)1 << 31);
  int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
  if (vec4) reflect_fold_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(gpad, gx, B, H, W, C, ph, pw);
  else reflect_fold_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(gpad, gx, B, H, W, C, ph, pw);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
