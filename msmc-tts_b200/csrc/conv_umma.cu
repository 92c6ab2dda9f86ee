// Implicit-GEMM convolution on the 5th-gen tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM.
//
//   dst[m, n] = dst_xf( sum_{tap, c} src_xf(src)[gather(m, tap), c] * W[tap][c][n] + bias[n] ) + residual[m, n]
//
// One CTA owns a 128 x BN output tile (M = flattened (b, hd, wd) output positions, N = output channels) and runs a
// 4-stage mbarrier pipeline over K = taps x (Cs/32) chunks of 32 tf32 (= one 128-byte swizzle row):
//   warps 0-3  A producers: thread t owns output row t; it gathers the 128 contiguous bytes of its source pixel for
//              the stage's tap/chunk from global memory (zero / reflect padding and the operand transform -- leaky
//              ReLU, relu'/tanh' masks -- are applied in registers: this is the part TMA cannot do), and stores them
//              in the canonical K-major SWIZZLE_128B layout, then fence.proxy.async + mbarrier arrive.
//              Thread 0 also issues ONE cp.async.bulk (TMA unit, UBLKCP) for the stage's weight tile: the weights
//              are pre-arranged in global memory as ready-to-use swizzled tile images (msmc_weight_image).
//   warp 4     single elected thread issues 4 x tcgen05.mma (M=128, N=BN, K=8) per stage and tcgen05.commit's the
//              stage back to the producers; after the last stage commits to the epilogue barrier.  Owns TMEM alloc.
//   warps 0-3  epilogue: tcgen05.ld 32 lanes x BN columns -> registers -> bias / activation / residual -> global.
// Precision: SPLIT = 3xTF32.  Each fp32 operand is split in registers into hi = top 19 bits (exact TF32) and
// lo = x - hi (exact in fp32), both tiles are staged, and every K-step issues three MMAs into the same TMEM
// accumulator: lo*hi + hi*lo + hi*hi (the dropped lo*lo term is 2^-22 relative).  The result matches an fp32 FMA
// chain to ~1e-6, so the fp32 parity tolerances and the bit-exact VQ indices downstream hold while the contraction
// runs on the tensor cores; the layers on this path are bandwidth/latency-bound, so the 3x MMA count is not the
// limiter.  SPLIT = false is plain TF32 (what the reference's cuDNN convolutions do on Ampere+ by default).
#include "umma.cuh"
#include <algorithm>
#include <cstdlib>

namespace msmc {
namespace {

// unsigned division by a runtime constant for dividends < 2^31: q = (umulhi(n, mul) + n) >> shr
struct FastDiv {
  uint32_t mul, shr, d;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.shr = l;
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.mul) + n) >> f.shr; }

struct UmmaArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* wimg;      // [tap][Cs/32][n_tile][BN rows x 128 B, swizzled]
  const float* bias;
  const float* residual;
  const float* dst_aux;
  float* dst;
};

__device__ __forceinline__ int reflect1(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}


// operand-transform classes the producer is specialised on (keeps the per-element code straight-line)
constexpr int XFC_NONE = 0, XFC_LRELU = 1, XFC_GENERIC = 2;

template <int XFC>
__device__ __forceinline__ float4 xf4(const msmc_conv_geom& g, float4 x, float4 ax) {
  if (XFC == XFC_LRELU) {
    // slope in (0, 1): leaky_relu(x) == max(x, slope * x)
    const float sl = g.src_slope;
    x.x = fmaxf(x.x, sl * x.x); x.y = fmaxf(x.y, sl * x.y); x.z = fmaxf(x.z, sl * x.z); x.w = fmaxf(x.w, sl * x.w);
  } else if (XFC == XFC_GENERIC) {
    x.x = apply_xf(g.src_xf, g.src_slope, x.x, ax.x);
    x.y = apply_xf(g.src_xf, g.src_slope, x.y, ax.y);
    x.z = apply_xf(g.src_xf, g.src_slope, x.z, ax.z);
    x.w = apply_xf(g.src_xf, g.src_slope, x.w, ax.w);
  }
  return x;
}

// fewest source channels the tensor-core kernels take: below 32 the single 32-channel chunk is ragged (zero-filled
// in the operand tile and the weight image), which wastes MMA width the CUDA-core fallbacks do not have to spare
constexpr int UM_MIN_CS = 8;
constexpr int UMF_PRODUCERS = 256;             // 8 producer / epilogue warps
constexpr int UMF_THREADS = UMF_PRODUCERS + 32;  // + the MMA warp

constexpr int UM_MAX_PHASE_TAPS = 256;

// Producer / epilogue warps: 8, or 16 for the one-CTA-per-SM configuration (BN = 128, three stages): with 9 warps
// on the SM every issue waited on a load or a dependent result and the tensor pipe idled (same finding as in the
// weight-gradient kernel); 16 warps take two rows per thread and stage instead of four.
template <int BN, int STAGES>
struct UmmaFwdCfg {
  static constexpr int PW = (BN == 128 && STAGES == 3) ? 16 : 8;
  static constexpr int RPT = UM_BM / (PW * 4);     // rows per thread and stage: 4 or 2
};
template <int BN, bool SPLIT, int STAGES, int XFC, bool TRANSPOSED>
__global__ void __launch_bounds__(UmmaFwdCfg<BN, STAGES>::PW * 32 + 32, (STAGES == 2 ? 2 : 1))
conv_umma_kernel(const UmmaArgs a) {
  constexpr int PW = UmmaFwdCfg<BN, STAGES>::PW;
  constexpr int RPT = UmmaFwdCfg<BN, STAGES>::RPT;
  constexpr int RSTEP = PW * 4;                // row slots per pass: a warp instruction covers 4 rows
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: A stages (16 KB per plane), B stages (BN*128 B per plane), barriers, tmem slot
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NP = SPLIT ? 2 : 1;             // planes per operand: hi (, lo)
  constexpr int A_PLANE = UM_BM * 128;
  constexpr int B_PLANE = BN * 128;
  constexpr int A_BYTES = NP * A_PLANE;
  constexpr int B_BYTES = NP * B_PLANE;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  // conv-transpose form: blockIdx.z is the phase (hd % sh, wd % sw); its destination rows form a regular
  // sub-grid and only the taps congruent with the phase contribute (no wasted taps)
  __shared__ short s_tap_kh[UM_MAX_PHASE_TAPS];
  __shared__ short s_tap_kw[UM_MAX_PHASE_TAPS];
  __shared__ int s_ntaps;
  int prh = 0, prw = 0, step_h = 1, step_w = 1;
  if (TRANSPOSED) {
    prh = blockIdx.z / g.sw;
    prw = blockIdx.z % g.sw;
    step_h = g.sh;
    step_w = g.sw;
  }
  const int Hp = TRANSPOSED ? (g.Hd - prh + step_h - 1) / step_h : g.Hd;
  const int Wp = TRANSPOSED ? (g.Wd - prw + step_w - 1) / step_w : g.Wd;
  const int64_t M = (int64_t)g.B * Hp * Wp;
  const int64_t m0 = (int64_t)blockIdx.x * UM_BM;
  if (m0 >= M) return;                       // (uniform per CTA; phases have slightly different row counts)
  const int n_tile = blockIdx.y;
  const int n_tiles = gridDim.y;
  const int KC = (g.Cs + UM_BK - 1) / UM_BK;      // the last chunk may be ragged (Cs % 4 == 0): zero-filled
  int T = g.KH * g.KW;
  if (TRANSPOSED) {
    if (tid == 0) {
      auto gcd = [](int x, int y) { while (y) { int t = x % y; x = y; y = t; } return x; };
      auto first_valid = [](int r, int p, int d, int s, int period, int K) {
        for (int k = 0; k < period && k < K; ++k) {
          int t = r + p - k * d;
          if (((t % s) + s) % s == 0) return k;
        }
        return K;
      };
      const int per_h = g.sh / gcd(g.dh % g.sh == 0 ? g.sh : g.dh % g.sh, g.sh);
      const int per_w = g.sw / gcd(g.dw % g.sw == 0 ? g.sw : g.dw % g.sw, g.sw);
      const int kh0 = first_valid(prh, g.ph, g.dh, g.sh, per_h, g.KH);
      const int kw0 = first_valid(prw, g.pw, g.dw, g.sw, per_w, g.KW);
      int c = 0;
      for (int kh = kh0; kh < g.KH; kh += per_h)
        for (int kw = kw0; kw < g.KW && c < UM_MAX_PHASE_TAPS; kw += per_w) {
          s_tap_kh[c] = (short)kh;
          s_tap_kw[c] = (short)kw;
          ++c;
        }
      s_ntaps = c;
    }
    __syncthreads();
    T = s_ntaps;
  }
  const int n_k = T * KC;
  constexpr int MMA_WARP = PW;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], PW + 1);   // one arrival per producer warp + 1 arrive.expect_tx (weights)
      mbar_init(&empty_bar[s], 1);                  // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    // allocate BN TMEM columns (power of two >= 32); the address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform -> uniform register

  if (warp < MMA_WARP) {
    // ================================= A producers =================================
    // Coalesced mapping: one warp instruction reads 4 output rows x 8 sixteen-byte chunks (4 full 128 B lines);
    // thread t owns chunk (t & 7) of rows (t >> 3) + 32*i, i = 0..3.
    const int chunk = tid & 7;
    const int rsub = tid >> 3;                 // 0..31
    const int r8 = rsub & 7;                   // (row & 7) is the same for all rows of this thread
    int pixb[RPT], hs0[RPT], ws0[RPT];               // batch pixel base / top-left source coordinate of each row
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int64_t m = m0 + rsub + RSTEP * i;
      if (m < M) {
        const int b = (int)(m / ((int64_t)Hp * Wp));
        const int rem = (int)(m % ((int64_t)Hp * Wp));
        const int hd = prh + (rem / Wp) * step_h, wd = prw + (rem % Wp) * step_w;
        pixb[i] = b * g.Hs * g.Ws;
        hs0[i] = TRANSPOSED ? hd + g.ph : hd * g.sh - g.ph;
        ws0[i] = TRANSPOSED ? wd + g.pw : wd * g.sw - g.pw;
      } else {
        pixb[i] = -1; hs0[i] = 0; ws0[i] = 0;
      }
    }
    constexpr bool NEED_AUX = (XFC == XFC_GENERIC);
    const bool need_aux = NEED_AUX && xf_needs_aux(g.src_xf);
    float4 v0[RPT], u0[RPT], v1[RPT], u1[RPT];          // two stages of loads in flight (prefetch distance 2)
    int kc_n = 0, kh_n = 0, kw_n = 0, ti_n = 0;   // (chunk, tap) of the NEXT stage to gather
    if (TRANSPOSED && T > 0) { kh_n = s_tap_kh[0]; kw_n = s_tap_kw[0]; }

    // raw global loads only (no dependent math), so they stay in flight across the barrier round-trips
    auto gather = [&](float4 (&v)[RPT], float4 (&u)[RPT]) {
      const int coff = kc_n * UM_BK + chunk * 4;
      const int dh = kh_n * g.dh, dw = kw_n * g.dw;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        int hs, ws;
        bool ok = pixb[i] >= 0;
        if (TRANSPOSED) {
          // the phase guarantees divisibility; only the range has to be checked
          const int th = hs0[i] - dh, tw = ws0[i] - dw;
          hs = th / g.sh;
          ws = tw / g.sw;
          ok = ok && th >= 0 && tw >= 0 && hs < g.Hs && ws < g.Ws;
        } else {
          hs = hs0[i] + dh;
          ws = ws0[i] + dw;
          if (g.pad_reflect) {
            hs = reflect1(hs, g.Hs);
            ws = reflect1(ws, g.Ws);
          } else {
            ok = ok && (unsigned)hs < (unsigned)g.Hs && (unsigned)ws < (unsigned)g.Ws;
          }
        }
        if (ok && coff < g.Cs) {
          const int pix = pixb[i] + hs * g.Ws + ws;
          v[i] = __ldg(reinterpret_cast<const float4*>(a.src + (int64_t)pix * g.ld_src + coff));
          if (NEED_AUX && need_aux)
            u[i] = __ldg(reinterpret_cast<const float4*>(a.src_aux + (int64_t)pix * g.ld_saux + coff));
        } else {
          // padding: every operand transform maps 0 -> 0, so zeros pass through the store path unchanged
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (NEED_AUX) u[i] = v[i];
        }
      }
      if (++kc_n == KC) {
        kc_n = 0;
        if (TRANSPOSED) {
          ++ti_n;
          if (ti_n < T) { kh_n = s_tap_kh[ti_n]; kw_n = s_tap_kw[ti_n]; }
        } else if (++kw_n == g.KW) { kw_n = 0; ++kh_n; }
      }
    };

    int s = 0;
    uint32_t ph = 0;
    const uint32_t dst_off = (uint32_t)(rsub >> 3) * 1024u + (uint32_t)r8 * 128u + (uint32_t)((chunk ^ r8) << 4);
    auto produce = [&](int ks, float4 (&v)[RPT], float4 (&u)[RPT]) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
        int64_t wk = ks;                                   // (tap * KC + chunk) of this stage
        if (TRANSPOSED) {
          const int ti = ks / KC, kc = ks - ti * KC;
          wk = (int64_t)(s_tap_kh[ti] * g.KW + s_tap_kw[ti]) * KC + kc;
        }
        const float* wsrc = a.wimg + (wk * n_tiles + n_tile) * (B_BYTES / 4);
        bulk_g2s(sB + s * B_BYTES, wsrc, B_BYTES, &full_bar[s]);
      }
      uint8_t* dstbase = sA + s * A_BYTES + dst_off;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float4 x = xf4<XFC>(g, v[i], NEED_AUX ? u[i] : make_float4(0.f, 0.f, 0.f, 0.f));
        uint8_t* d = dstbase + i * (RSTEP * 128);   // rows advance by RSTEP -> RSTEP / 8 one-KB swizzle atoms
        if (SPLIT) {
          const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
          *reinterpret_cast<float4*>(d) = hi;
          *reinterpret_cast<float4*>(d + A_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
        } else {
          *reinterpret_cast<float4*>(d) = x;
        }
      }
      if (ks + 2 < n_k) gather(v, u);   // refill this register set with the loads of stage ks + 2
      publish_and_arrive_warp(&full_bar[s]);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    };

    if (n_k > 0) gather(v0, u0);
    if (n_k > 1) gather(v1, u1);
    for (int ks = 0; ks < n_k; ks += 2) {
      produce(ks, v0, u0);
      if (ks + 1 < n_k) produce(ks + 1, v1, u1);
    }

    // ================================= epilogue =================================
    // TMEM lane == accumulator row; warps w and w+4 share lanes 32*(w%4).. and split the columns
    const int lane_grp = warp & 3;
    const int64_t mt = m0 + lane_grp * 32 + (tid & 31);      // row inside this phase
    const bool row_ok = mt < M;
    int64_t m = mt;                                          // destination pixel index
    if (TRANSPOSED && row_ok) {
      const int b = (int)(mt / ((int64_t)Hp * Wp));
      const int rem = (int)(mt % ((int64_t)Hp * Wp));
      m = ((int64_t)b * g.Hd + prh + (rem / Wp) * step_h) * g.Wd + prw + (rem % Wp) * step_w;
    }
    if (n_k == 0) {
      // a phase without taps still owes bias / zeros to its destination rows
      if (row_ok) {
        const int n0z = n_tile * BN;
        for (int j = (warp >> 2) * (BN / (PW / 4)); j < (warp >> 2) * (BN / (PW / 4)) + BN / (PW / 4); ++j) {
          const int n = n0z + j;
          if (n < g.Cd) {
            float x = a.bias ? __ldg(a.bias + n) : 0.f;
            if (g.dst_xf != MSMC_XF_NONE)
              x = apply_xf(g.dst_xf, g.dst_slope, x, xf_needs_aux(g.dst_xf) ? __ldg(a.dst_aux + m * g.ld_daux + n) : 0.f);
            if (a.residual) x += __ldg(a.residual + m * g.ld_res + n);
            a.dst[m * g.ld_dst + n] = x;
          }
        }
      }
    } else {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    const int n0 = n_tile * BN;
    const bool dneed_aux = xf_needs_aux(g.dst_xf);
    constexpr int CHALF = BN / (PW / 4);         // the PW / 4 warps of a TMEM lane quadrant split the columns
    const int cbeg = (warp >> 2) * CHALF;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CHALF; c0 += 16) {
      float acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      if (row_ok) {
        const bool full16 = n0 + c0 + 16 <= g.Cd;
        if (full16 && (!dneed_aux || ((g.ld_daux & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dst_aux) & 15) == 0)) && (g.ld_dst & 3) == 0 && (!a.residual || (g.ld_res & 3) == 0)) {
          // fast path: 16 full columns, vector loads of bias / residual, vector stores
          float* out = a.dst + m * g.ld_dst + n0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 x = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            if (a.bias) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + c0 + j));
              x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
            }
            if (g.dst_xf != MSMC_XF_NONE) {
              float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
              if (dneed_aux) av = __ldg(reinterpret_cast<const float4*>(a.dst_aux + m * g.ld_daux + n0 + c0 + j));
              x.x = apply_xf(g.dst_xf, g.dst_slope, x.x, av.x); x.y = apply_xf(g.dst_xf, g.dst_slope, x.y, av.y);
              x.z = apply_xf(g.dst_xf, g.dst_slope, x.z, av.z); x.w = apply_xf(g.dst_xf, g.dst_slope, x.w, av.w);
            }
            if (a.residual) {
              const float4 rv = __ldg(reinterpret_cast<const float4*>(a.residual + m * g.ld_res + n0 + c0 + j));
              x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
            }
            *reinterpret_cast<float4*>(out + j) = x;
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c0 + j;
          if (n < g.Cd) {
            float x = acc[j];
            if (a.bias) x += __ldg(a.bias + n);
            if (g.dst_xf != MSMC_XF_NONE) {
              const float aux = dneed_aux ? __ldg(a.dst_aux + m * g.ld_daux + n) : 0.f;
              x = apply_xf(g.dst_xf, g.dst_slope, x, aux);
            }
            if (a.residual) x += __ldg(a.residual + m * g.ld_res + n);
            acc[j] = x;
          }
        }
        float* out = a.dst + m * g.ld_dst + n0 + c0;
        if (full16 && (g.ld_dst & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < g.Cd) out[j] = acc[j];
        }
      }
    }
    }
    tc_fence_before();
  } else {
    // ================================= MMA issuer =================================
    // instruction descriptor: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(UM_BM >> 4) << 24);
    // the whole warp walks the loop with warp-uniform values; one elected lane issues (predicated, straight-line:
    // see umma_tf32_pred -- inside `if (lane == 0)` every MMA cost ~90 cycles of issue latency)
    {
      const uint32_t elected = elect_one();
      int s = 0;
      uint32_t ph = 0;
      const uint32_t a_desc0 = desc_lo_k(smem_u32(sA)), b_desc0 = desc_lo_k(smem_u32(sB));
      for (int ks = 0; ks < n_k; ++ks) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t ad = a_desc0 + (uint32_t)s * (A_BYTES >> 4);
        const uint32_t bd = b_desc0 + (uint32_t)s * (B_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < UM_BK / 8; ++k) {
          // advance 8 tf32 = 32 bytes (2 descriptor units) along K inside the 128-byte swizzle row
          const uint32_t a_hi = ad + 2 * k, b_hi = bd + 2 * k;
          const uint32_t acc = (ks > 0 || k > 0) ? 1u : 0u;
          if (SPLIT) {
            const uint32_t a_lo = a_hi + (A_PLANE >> 4), b_lo = b_hi + (B_PLANE >> 4);
            umma_tf32_pred<DESC_HI_K>(tmem_base, a_lo, b_hi, IDESC, acc, elected);   // small terms first
            umma_tf32_pred<DESC_HI_K>(tmem_base, a_hi, b_lo, IDESC, 1u, elected);
            umma_tf32_pred<DESC_HI_K>(tmem_base, a_hi, b_hi, IDESC, 1u, elected);
          } else {
            umma_tf32_pred<DESC_HI_K>(tmem_base, a_hi, b_hi, IDESC, acc, elected);
          }
        }
        umma_commit_pred(&empty_bar[s], elected);   // frees the stage when these MMAs have read it
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (n_k > 0) umma_commit_pred(accum_bar, elected);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}


// ------------------------------------------------------------------------------------------------------------
// Stride-1 convolution with TAP REUSE.  For a stride-1 conv every tap reads the same source rows shifted by a
// constant number of pixels, so the operand tile of one 32-channel chunk is staged ONCE (128 + max-shift rows,
// transformed / split in registers as before) and each tap's tcgen05.mma simply starts its shared-memory
// descriptor `shift` rows further down: start address += shift * 128 B with the descriptor's matrix-base-offset
// field = (shift & 7), which tells the tensor core where in the 8-row SWIZZLE_128B pattern the tile begins.
// Producer work drops by the number of taps (3..11x on the MRF / FFN / MPD convs) and the kernel becomes
// MMA / HBM bound instead of producer bound.  A dedicated warp streams the weight tiles (cp.async.bulk) through
// their own ring, decoupled from the operand ring.
//   applies to: forward form, stride 1, zero padding, 1-D tap pattern in flattened pixel space
//               (Hs == 1, or KW == 1 with pw == 0), max shift <= 64 rows.
// ------------------------------------------------------------------------------------------------------------
constexpr int RU_ROWS = 192;                  // 128 output rows + up to 64 rows of tap reach
constexpr int RU_PRODUCERS = 256;
constexpr int RU_THREADS = RU_PRODUCERS + 64;  // + MMA warp + weight-loader warp

// ---- thread-block-cluster helpers (split-K over the source channels: the CTAs of a cluster share one output tile)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta_rank(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}

struct ReuseArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* wimg;
  const float* bias;
  const float* residual;
  const float* dst_aux;
  float* dst;
  int Ls, Ld;            // pixels per batch element (source / destination)
  int pad_rows;          // ph * Ws + pw
  int tap_stride;        // rows between consecutive taps (dh * Ws or dw)
  int n_taps;
  int tiles_per_batch;
  int k_splits;          // cluster (1, 1, k_splits): CTA z of a cluster takes the z-th slice of the channel chunks
};


// Producer / epilogue warps: 8, or 16 for the one-CTA-per-SM configuration (BN = 128, two operand stages).
template <int BN, int NA>
struct ReuseCfg {
  static constexpr int PW = (BN == 128 && NA == 2) ? 16 : 8;
  static constexpr int THREADS = PW * 32 + 64;         // + MMA warp + weight-loader warp
};
template <int BN, bool SPLIT, int NA, int NBS, int XFC>
__global__ void __launch_bounds__(ReuseCfg<BN, NA>::THREADS, (NA == 1 ? 2 : 1))
conv_umma_reuse_kernel(const ReuseArgs a) {
  constexpr int PW = ReuseCfg<BN, NA>::PW;
  constexpr int RSTEP = PW * 4;                  // row slots per pass (a warp instruction covers 4 rows)
  constexpr int RPT = RU_ROWS / RSTEP;           // rows per thread and chunk: 6 or 3
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr int A_PLANE = RU_ROWS * 128;        // 24 KB
  constexpr int B_PLANE = BN * 128;
  constexpr int A_BYTES = NP * A_PLANE;
  constexpr int B_BYTES = NP * B_PLANE;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * A_BYTES;
  uint64_t* fa = reinterpret_cast<uint64_t*>(sB + NBS * B_BYTES);
  uint64_t* ea = fa + NA;
  uint64_t* fb = ea + NA;
  uint64_t* eb = fb + NBS;
  uint64_t* accum_bar = eb + NBS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / a.tiles_per_batch;
  const int l0 = (blockIdx.x - b * a.tiles_per_batch) * UM_BM;
  const int n_tile = blockIdx.y, n_tiles = gridDim.y;
  const int KC = (g.Cs + UM_BK - 1) / UM_BK;
  // split-K: launches with few output tiles and a long channel loop (FFN second conv: 60 tiles x 96 MMA steps) run
  // as clusters of KS CTAs per tile; CTA ks reduces chunks [kc_beg, kc_end), CTA 0 sums the accumulators (in rank
  // order, through its own shared memory) and runs the epilogue
  const int KS = a.k_splits, ks = blockIdx.z;
  const int kc_beg = (KC * ks) / KS, kc_end = (KC * (ks + 1)) / KS;
  const int T = a.n_taps;
  const int r_in = UM_BM + (T - 1) * a.tap_stride;      // rows staged per chunk (<= RU_ROWS)
  constexpr int MMA_WARP = PW;                  // warp MMA_WARP + 1 streams the weight tiles

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&fa[i], PW); mbar_init(&ea[i], 1); }
    for (int i = 0; i < NBS; ++i) { mbar_init(&fb[i], 1); mbar_init(&eb[i], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform -> uniform register

  if (warp < MMA_WARP) {
    // ================================= operand producers =================================
    const int chunk = tid & 7;
    const int rsub = tid >> 3;                 // 0..RSTEP-1; rows rsub + RSTEP*i, i = 0..RPT-1
    const int r8 = rsub & 7;
    constexpr bool NEED_AUX = (XFC == XFC_GENERIC);
    const bool need_aux = NEED_AUX && xf_needs_aux(g.src_xf);
    const int64_t pix0 = (int64_t)b * a.Ls;
    float4 v[RPT], u[RPT];
    auto gather = [&](int kc) {
      const int coff = kc * UM_BK + chunk * 4;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rsub + RSTEP * i;
        const int p = l0 - a.pad_rows + r;               // source pixel inside this batch element
        if (r < r_in && (unsigned)p < (unsigned)a.Ls && coff < g.Cs) {
          v[i] = __ldg(reinterpret_cast<const float4*>(a.src + (pix0 + p) * g.ld_src + coff));
          if (NEED_AUX && need_aux)
            u[i] = __ldg(reinterpret_cast<const float4*>(a.src_aux + (pix0 + p) * g.ld_saux + coff));
        } else {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (NEED_AUX) u[i] = v[i];
        }
      }
    };
    gather(kc_beg);
    int sa = 0;
    uint32_t pa = 0;
    const uint32_t dst_off = (uint32_t)(rsub >> 3) * 1024u + (uint32_t)r8 * 128u + (uint32_t)((chunk ^ r8) << 4);
    for (int kc = kc_beg; kc < kc_end; ++kc) {
      mbar_wait(&ea[sa], pa ^ 1u);
      uint8_t* dstbase = sA + sa * A_BYTES + dst_off;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        if (rsub + RSTEP * i < r_in) {
          const float4 x = xf4<XFC>(g, v[i], NEED_AUX ? u[i] : make_float4(0.f, 0.f, 0.f, 0.f));
          uint8_t* d = dstbase + i * (RSTEP * 128);
          if (SPLIT) {
            const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
            *reinterpret_cast<float4*>(d) = hi;
            *reinterpret_cast<float4*>(d + A_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
          } else {
            *reinterpret_cast<float4*>(d) = x;
          }
        }
      }
      if (kc + 1 < kc_end) gather(kc + 1);
      publish_and_arrive_warp(&fa[sa]);
      if (++sa == NA) { sa = 0; pa ^= 1u; }
    }
    mbar_wait(accum_bar, 0);          // this CTA's MMAs are complete (its operand stages are dead)
    tc_fence_after();
  } else if (warp == MMA_WARP) {
    // ================================= MMA issuer =================================
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(UM_BM >> 4) << 24);
    {
      const uint32_t elected = elect_one();     // convergent, predicated issue (see umma_tf32_pred)
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const uint32_t a_desc0 = desc_lo_k(smem_u32(sA)), b_desc0 = desc_lo_k(smem_u32(sB));
      for (int kc = kc_beg; kc < kc_end; ++kc) {
        mbar_wait(&fa[sa], pa);
        // tap = row shift of the staged tile: the descriptor start moves by tap_stride rows of 128 B (8 units); the
        // swizzle is a function of the absolute shared-memory address, so the base-offset field stays 0
        uint32_t ad = a_desc0 + (uint32_t)sa * (A_BYTES >> 4);
        const uint32_t a_step = (uint32_t)a.tap_stride * 8u;
        for (int t = 0; t < T; ++t, ad += a_step) {
          mbar_wait(&fb[sb], pb);
          tc_fence_after();
          const uint32_t bd = b_desc0 + (uint32_t)sb * (B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < UM_BK / 8; ++k) {
            const uint32_t a_hi = ad + 2 * k, b_hi = bd + 2 * k;
            const uint32_t acc = (kc > kc_beg || t > 0 || k > 0) ? 1u : 0u;
            if (SPLIT) {
              const uint32_t a_lo = a_hi + (A_PLANE >> 4), b_lo = b_hi + (B_PLANE >> 4);
              umma_tf32_pred<DESC_HI_K>(tmem_base, a_lo, b_hi, IDESC, acc, elected);
              umma_tf32_pred<DESC_HI_K>(tmem_base, a_hi, b_lo, IDESC, 1u, elected);
              umma_tf32_pred<DESC_HI_K>(tmem_base, a_hi, b_hi, IDESC, 1u, elected);
            } else {
              umma_tf32_pred<DESC_HI_K>(tmem_base, a_hi, b_hi, IDESC, acc, elected);
            }
          }
          umma_commit_pred(&eb[sb], elected);
          if (++sb == NBS) { sb = 0; pb ^= 1u; }
        }
        umma_commit_pred(&ea[sa], elected);
        if (++sa == NA) { sa = 0; pa ^= 1u; }
      }
      umma_commit_pred(accum_bar, elected);
    }
    __syncwarp();
  } else {
    // ================================= weight-tile loader =================================
    if ((tid & 31) == 0) {
      int sb = 0;
      uint32_t pb = 0;
      for (int kc = kc_beg; kc < kc_end; ++kc)
        for (int t = 0; t < T; ++t) {
          mbar_wait(&eb[sb], pb ^ 1u);
          mbar_arrive_expect_tx(&fb[sb], B_BYTES);
          const float* wsrc = a.wimg + (((int64_t)t * KC + kc) * n_tiles + n_tile) * (B_BYTES / 4);
          bulk_g2s(sB + sb * B_BYTES, wsrc, B_BYTES, &fb[sb]);
          if (++sb == NBS) { sb = 0; pb ^= 1u; }
        }
    }
    __syncwarp();
  }

  // ============ split-K: the accumulators of CTAs 1 .. KS-1 travel to CTA 0's (now dead) operand stages ============
  // layout per partial: [column][128 rows] fp32 -- a warp's 32 rows are one 128-byte line on both sides
  const int lane_grp = warp & 3;
  const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
  constexpr int CHALF = BN / (PW / 4);           // the PW / 4 warps of a TMEM lane quadrant split the columns
  const int cbeg = (warp >> 2) * CHALF;
  const int erow = lane_grp * 32 + (tid & 31);      // accumulator row of this thread (epilogue warps)
  if (KS > 1) {
    cluster_sync_all();               // every CTA of the cluster has finished its MMAs
    if (ks > 0 && warp < MMA_WARP) {
      const uint32_t dst0 = map_to_cta_rank(smem_u32(smem) + (uint32_t)(ks - 1) * (uint32_t)(BN * UM_BM * 4), 0u) +
                            (uint32_t)erow * 4u;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + CHALF; c0 += 16) {
        float acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst0 + (uint32_t)((c0 + j) * UM_BM * 4)), "f"(acc[j])
                       : "memory");
      }
      tc_fence_before();
    }
    cluster_sync_all();               // the partials have landed (release / acquire at cluster scope)
  }

  if (warp < MMA_WARP && ks == 0) {
    // ================================= epilogue =================================
    const int l = l0 + erow;
    const bool row_ok = l < a.Ld;
    const int64_t m = (int64_t)b * a.Ld + l;
    const int n0 = n_tile * BN;
    const bool dneed_aux = xf_needs_aux(g.dst_xf);
    const float* part = reinterpret_cast<const float*>(smem) + erow;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CHALF; c0 += 16) {
      float acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      for (int p = 0; p < KS - 1; ++p)              // rank order: (acc0 + acc1) + acc2 ...
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += part[(p * BN + c0 + j) * UM_BM];
      if (row_ok) {
        const bool full16 = n0 + c0 + 16 <= g.Cd;
        if (full16 && (!dneed_aux || ((g.ld_daux & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dst_aux) & 15) == 0)) && (g.ld_dst & 3) == 0 && (!a.residual || (g.ld_res & 3) == 0)) {
          // fast path: 16 full columns, vector loads of bias / residual, vector stores
          float* out = a.dst + m * g.ld_dst + n0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 x = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            if (a.bias) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + c0 + j));
              x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
            }
            if (g.dst_xf != MSMC_XF_NONE) {
              float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
              if (dneed_aux) av = __ldg(reinterpret_cast<const float4*>(a.dst_aux + m * g.ld_daux + n0 + c0 + j));
              x.x = apply_xf(g.dst_xf, g.dst_slope, x.x, av.x); x.y = apply_xf(g.dst_xf, g.dst_slope, x.y, av.y);
              x.z = apply_xf(g.dst_xf, g.dst_slope, x.z, av.z); x.w = apply_xf(g.dst_xf, g.dst_slope, x.w, av.w);
            }
            if (a.residual) {
              const float4 rv = __ldg(reinterpret_cast<const float4*>(a.residual + m * g.ld_res + n0 + c0 + j));
              x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
            }
            *reinterpret_cast<float4*>(out + j) = x;
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c0 + j;
          if (n < g.Cd) {
            float x = acc[j];
            if (a.bias) x += __ldg(a.bias + n);
            if (g.dst_xf != MSMC_XF_NONE) {
              const float aux = dneed_aux ? __ldg(a.dst_aux + m * g.ld_daux + n) : 0.f;
              x = apply_xf(g.dst_xf, g.dst_slope, x, aux);
            }
            if (a.residual) x += __ldg(a.residual + m * g.ld_res + n);
            acc[j] = x;
          }
        }
        float* out = a.dst + m * g.ld_dst + n0 + c0;
        if (full16 && (g.ld_dst & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < g.Cd) out[j] = acc[j];
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  dW[(tap, cs), cd] = sum_m  xf(src)[gather(m, tap), cs] * xf(gout)[m, cd]
//
// The reduction runs over output positions m, which are the OUTER index of both channels-last operands, so both
// UMMA operands are MN-major: a 128-byte shared-memory row is "32 consecutive channels of one position", four
// consecutive positions form one SWIZZLE_128B_BASE32B atom (the only MN-major layout TF32 has; one tcgen05.mma
// with K=8 spans two atoms, SBO = 512 B), 32-channel blocks sit LBO = 4 KB apart.  One CTA owns 4 row-blocks of dW (a row-block = 32 channels of one tap -> M = 128 accumulator lanes) x BN
// output channels and loops over its slice of positions, 32 per stage.  Producer thread (blk = warp, p = lane)
// gathers one 128-byte row per operand per stage; lanes are positions, so the bias gradient (column sums of gout)
// is a warp-shuffle reduction.  Partial sums per position-slice go to the workspace and are reduced by the same
// deterministic second pass as the CUDA-core path.
// ------------------------------------------------------------------------------------------------------------
struct UmmaWgradArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* gout;
  const float* gout_aux;
  float* partial;          // [splits][Ktot*Cd + Cd]
  int64_t rows_per_split;  // multiple of 32
  int want_bias;
  int gvec;                // gout rows can be read with 16-byte loads
  FastDiv div_hw, div_w;   // m -> (b, rem) by Hd*Wd, rem -> (hd, wd) by Wd
  int dry;                 // bring-up/profiling switch: 1 = producers skip loads and stores (MMA pipeline only),
                           //                            2 = producers load and store but the MMA warp issues nothing
};

// (A K-major variant -- producers transposing both operands with warp shuffles -- was measured slower and removed.)
// Producer warps: 8 (two CTAs per SM at BN <= 64) or 16 at BN = 128 (one CTA per SM: with 8 producer warps the SM
// held 9 warps whose every issue waited ~8 cycles -- long scoreboard 3.0, wait 2.0, instruction fetch 1.1 -- and the
// tensor pipe sat at 15 %; two warps per operand block, 16 positions each, halve the registers per thread and double
// the loads in flight).
template <int BN, bool SPLIT, int STAGES, int XFC>
__global__ void __launch_bounds__((BN == 128 ? 16 : 8) * 32 + 32, (BN <= 64 ? 2 : 1))
conv_wgrad_umma_kernel(const UmmaWgradArgs a) {
  constexpr int PW = (BN == 128 ? 16 : 8);     // producer / epilogue warps
  constexpr int PH = PW / 8;                   // warps per operand block (each takes 32 / PH positions of a stage)
  constexpr int NR = 8 / PH;                   // 16-byte loads per thread, operand and stage
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr int A_PLANE = 4 * 4096;            // 4 row-blocks x 32 positions x 128 B
  constexpr int NB = BN / 32;                  // 32-channel blocks of the gout operand
  constexpr int B_PLANE = NB * 4096;
  constexpr int A_BYTES = NP * A_PLANE;
  constexpr int B_BYTES = NP * B_PLANE;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = (g.Cs + 31) / 32;                     // the last channel block may be ragged (Cs % 4 == 0)
  const int RB = g.KH * g.KW * KC;                     // row-blocks of dW
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int rb0 = blockIdx.x * 4;
  const int n0 = blockIdx.y * BN;
  const int64_t mbeg = (int64_t)blockIdx.z * a.rows_per_split;
  const int64_t mend = min(M, mbeg + a.rows_per_split);
  const int n_k = (int)((mend - mbeg + 31) / 32);      // stages of 32 positions (>= 1 by construction)
  constexpr int MMA_WARP = PW;
  __shared__ float s_bias[4][2][32];          // per-half bias partials when two warps share an operand block

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], PW);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform -> uniform register

  if (warp < MMA_WARP) {
    // ======= producers: the first PW/2 warps stage the source operand, the others the output-gradient operand =======
    // PH warps per 32-channel block; inside the warp one instruction reads 4 positions x 128 B:
    // thread owns 16-byte chunk (lane & 7) of positions half*(32/PH) + (lane >> 3) + 4*i, i = 0..NR-1.
    const bool is_a = warp < PW / 2;
    const int wsub = warp % (PW / 2);
    const int blk = wsub / PH;                 // 32-channel block of this warp's operand
    const int half = wsub % PH;                // which 32 / PH positions of every stage
    const int chunk = lane & 7;
    const int psub = lane >> 3;                // 0..3
    // A side: which (tap, channel block) this warp stages
    const int rb = rb0 + blk;
    const bool rb_ok = rb < RB;
    int tap = 0, cb = 0, kh = 0, kw = 0;
    if (rb_ok) { tap = rb / KC; cb = rb - tap * KC; kh = tap / g.KW; kw = tap - kh * g.KW; }
    // B side
    const bool nb_ok = blk < NB;
    const int nbase = n0 + blk * 32;
    const bool gvec_ok = a.gvec && (nbase + 32 <= g.Cd);
    const bool gneed_aux = xf_needs_aux(g.dst_xf);
    constexpr bool NEED_AUX = (XFC == XFC_GENERIC);
    const bool need_aux = NEED_AUX && xf_needs_aux(g.src_xf);
    const bool do_bias = a.want_bias && blockIdx.x == 0 && !is_a && nb_ok;
    const bool active = is_a ? true : nb_ok;
    // one stage of loads in flight per thread (two stages cost 128 registers here and spilled to local memory:
    // the ncu capture in profiles/ showed 400 MB of DRAM writes from the spill traffic alone)
    float4 v[NR], u[NR];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    int64_t m_next = mbeg + half * (4 * NR) + psub;     // position of row i = 0 of the next stage to gather
    const bool small_w = g.Wd < 4;             // carry-propagating decode below needs Wd >= 4

    auto gather = [&]() {
      // decode the first position with two multiply-shift divisions, walk the other seven by +4 with carries
      int b = 0, hd = 0, wd = 0;
      if (is_a && m_next < mend) {
        const uint32_t mu = (uint32_t)m_next;
        const uint32_t bb = fdiv(mu, a.div_hw);
        const uint32_t rem = mu - bb * a.div_hw.d;
        const uint32_t hh = fdiv(rem, a.div_w);
        b = (int)bb; hd = (int)hh; wd = (int)(rem - hh * a.div_w.d);
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int64_t m = m_next + 4 * i;
        const bool m_ok = m < mend;
        if (is_a) {
          bool ok = m_ok && rb_ok;
          int hs = hd * g.sh + kh * g.dh - g.ph;
          int ws = wd * g.sw + kw * g.dw - g.pw;
          if (g.pad_reflect) { hs = reflect1(hs, g.Hs); ws = reflect1(ws, g.Ws); }
          else ok = ok && (unsigned)hs < (unsigned)g.Hs && (unsigned)ws < (unsigned)g.Ws;
          if (ok && cb * 32 + chunk * 4 < g.Cs) {
            const int64_t pix = ((int64_t)b * g.Hs + hs) * g.Ws + ws;
            v[i] = __ldg(reinterpret_cast<const float4*>(a.src + pix * g.ld_src + cb * 32 + chunk * 4));
            if (NEED_AUX && need_aux)
              u[i] = __ldg(reinterpret_cast<const float4*>(a.src_aux + pix * g.ld_saux + cb * 32 + chunk * 4));
          } else {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NEED_AUX) u[i] = v[i];
          }
          // advance (b, hd, wd) by 4 positions
          if (small_w) {
            const uint32_t mu = (uint32_t)min(m + 4, mend - 1 > 0 ? mend - 1 : (int64_t)0);
            const uint32_t bb = fdiv(mu, a.div_hw);
            const uint32_t rem = mu - bb * a.div_hw.d;
            const uint32_t hh = fdiv(rem, a.div_w);
            b = (int)bb; hd = (int)hh; wd = (int)(rem - hh * a.div_w.d);
          } else {
            wd += 4;
            if (wd >= g.Wd) { wd -= g.Wd; if (++hd == g.Hd) { hd = 0; ++b; } }
          }
        } else if (nb_ok) {
          if (m_ok && gvec_ok) {
            v[i] = __ldg(reinterpret_cast<const float4*>(a.gout + m * g.ld_dst + nbase + chunk * 4));
            if (gneed_aux)
              u[i] = __ldg(reinterpret_cast<const float4*>(a.gout_aux + m * g.ld_daux + nbase + chunk * 4));
          } else {
            float t4[4], u4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nbase + chunk * 4 + j;
              const bool in = m_ok && n < g.Cd;
              t4[j] = in ? __ldg(a.gout + m * g.ld_dst + n) : 0.f;
              u4[j] = (in && gneed_aux) ? __ldg(a.gout_aux + m * g.ld_daux + n) : 0.f;
            }
            v[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
            u[i] = make_float4(u4[0], u4[1], u4[2], u4[3]);
          }
        }
      }
      m_next += 32;
    };

    if (a.dry != 1) gather();
    int s = 0;
    uint32_t ph = 0;
    for (int ks = 0; ks < n_k; ++ks) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      if (active && a.dry != 1) {
        uint8_t* base = (is_a ? sA + s * A_BYTES : sB + s * B_BYTES) + (uint32_t)blk * 4096u;
        constexpr int PLANE_A = A_PLANE, PLANE_B = B_PLANE;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const int p = half * (4 * NR) + psub + 4 * i;      // position within the stage = K index
          float4 x;
          if (is_a) {
            x = xf4<XFC>(g, v[i], NEED_AUX ? u[i] : make_float4(0.f, 0.f, 0.f, 0.f));
          } else {
            x = v[i];
            if (g.dst_xf != MSMC_XF_NONE) {
              const float4 ay = gneed_aux ? u[i] : make_float4(0.f, 0.f, 0.f, 0.f);
              x.x = apply_xf(g.dst_xf, g.dst_slope, x.x, ay.x);
              x.y = apply_xf(g.dst_xf, g.dst_slope, x.y, ay.y);
              x.z = apply_xf(g.dst_xf, g.dst_slope, x.z, ay.z);
              x.w = apply_xf(g.dst_xf, g.dst_slope, x.w, ay.w);
            }
            if (do_bias) { bsum[0] += x.x; bsum[1] += x.y; bsum[2] += x.z; bsum[3] += x.w; }
          }
          uint8_t* d = base + mn_off(p, chunk);
          if (SPLIT) {
            const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
            *reinterpret_cast<float4*>(d) = hi;
            *reinterpret_cast<float4*>(d + (is_a ? PLANE_A : PLANE_B)) =
                make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
          } else {
            *reinterpret_cast<float4*>(d) = x;
          }
        }
      }
      if (ks + 1 < n_k && a.dry != 1) gather();
      publish_and_arrive_warp(&full_bar[s]);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }

    // ================================= epilogue =================================
    float* part = a.partial + (int64_t)blockIdx.z * (Ktot * g.Cd + g.Cd);
    if (a.want_bias && blockIdx.x == 0) {        // (CTA-uniform)
      // lanes sharing a chunk differ in bits 3 and 4: fixed-order butterfly over the 4 position sub-indices; with
      // two warps per block the halves meet in shared memory and are added in half order
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = bsum[j];
          t += __shfl_xor_sync(0xffffffffu, t, 8);
          t += __shfl_xor_sync(0xffffffffu, t, 16);
          if (psub == 0) s_bias[blk][half][chunk * 4 + j] = t;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory");
      if (do_bias && half == 0 && psub == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = nbase + chunk * 4 + j;
          float t = s_bias[blk][0][chunk * 4 + j];
          if (PH == 2) t += s_bias[blk][1][chunk * 4 + j];
          if (n < g.Cd) part[Ktot * g.Cd + n] = t;
        }
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // TMEM lane = accumulator row = (row-block q, channel lane), q = warp % 4 (a warp reaches lane quadrant
    // warp % 4 only); the PW / 4 warps of a quadrant split the columns
    const int q = warp & 3;
    const int rb_e = rb0 + q;
    const bool rbe_ok = rb_e < RB;
    int tap_e = 0, cb_e = 0;
    if (rbe_ok) { tap_e = rb_e / KC; cb_e = rb_e - tap_e * KC; }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t krow = (int64_t)tap_e * g.Cs + cb_e * 32 + lane;    // row of dW this thread owns
    constexpr int CSL = BN / (PW / 4);
    const int cbeg = (warp >> 2) * CSL;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CSL; c0 += 16) {
      float acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      if (rbe_ok && cb_e * 32 + lane < g.Cs) {
        if (n0 + c0 + 16 <= g.Cd && (g.Cd & 3) == 0) {
          // (the workspace is 256-byte aligned and every partial slab is a multiple of Cd floats long)
          float* out = part + krow * g.Cd + n0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            if (n < g.Cd) part[krow * g.Cd + n] = acc[j];
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ================================= MMA issuer =================================
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((1u << 15) | (1u << 16)) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
    {
      const uint32_t elected = elect_one();     // convergent, predicated issue (see umma_tf32_pred)
      int s = 0;
      uint32_t ph = 0;
      const uint32_t a_desc0 = desc_lo_mn(smem_u32(sA));
      const uint32_t b_desc0 = desc_lo_mn(smem_u32(sB));
      for (int ks = 0; ks < n_k; ++ks) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t ad = a_desc0 + (uint32_t)s * (A_BYTES >> 4);
        const uint32_t bd = b_desc0 + (uint32_t)s * (B_BYTES >> 4);
        if (a.dry != 2) {
#pragma unroll
          for (int kg = 0; kg < 4; ++kg) {
            // one MMA = 8 positions: MN-major -> two 4-row atoms (1 KB = 64 units) per 32-channel block;
            //                        K-major  -> 32 bytes (2 units) further along every 128-byte channel row
            constexpr uint32_t KSTEP = 64u;
            constexpr uint32_t HI = DESC_HI_MN;
            const uint32_t a_hi = ad + kg * KSTEP, b_hi = bd + kg * KSTEP;
            const uint32_t acc = (ks > 0 || kg > 0) ? 1u : 0u;
            if (SPLIT) {
              const uint32_t a_lo = a_hi + (A_PLANE >> 4), b_lo = b_hi + (B_PLANE >> 4);
              umma_tf32_pred<HI>(tmem_base, a_lo, b_hi, IDESC, acc, elected);
              umma_tf32_pred<HI>(tmem_base, a_hi, b_lo, IDESC, 1u, elected);
              umma_tf32_pred<HI>(tmem_base, a_hi, b_hi, IDESC, 1u, elected);
            } else {
              umma_tf32_pred<HI>(tmem_base, a_hi, b_hi, IDESC, acc, elected);
            }
          }
        }
        umma_commit_pred(&empty_bar[s], elected);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit_pred(accum_bar, elected);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient of stride-1 convolutions with TAP REUSE.
//
// dW[tap][cs][cd] = sum_p src[p + tap*d - pad][cs] * gout[p][cd]: every tap reads the SAME source rows, shifted by
// tap*d positions.  In the MN-major operand layout a shared-memory row is one position, so
//   * the K extent (positions) of a tap starts `tap*d` rows further down the staged tile, and
//   * the four 32-lane M blocks of one tcgen05.mma are addressed through the descriptor's leading-dimension byte
//     offset (LBO).  Setting LBO = d*128 B makes the four M blocks FOUR CONSECUTIVE TAPS of one 32-channel block:
//     one MMA accumulates dW[4 taps][32 channels][BN] from a single staged copy of the source rows.
// The source tile (32 positions + tap reach) is therefore staged once per stage instead of once per tap and
// per row-block CTA (11 x fewer operand bytes on the k=11 MRF convs) and one CTA owns every tap of its channel
// blocks: accumulator group (cb, tap-group) lives in its own BN-column slice of TMEM (<= 512 columns).
//   applies to: forward-form stride-1 convs with a 1-D tap pattern in flattened position space (same rule as the
//   forward tap-reuse kernel), Cs % 32 == 0, Cd % 32 == 0, 3 <= taps, 32 + (4*ceil(taps/4) - 1) * d <= 96 rows.
// ------------------------------------------------------------------------------------------------------------
constexpr int WR_POS = 32;                 // positions (K) per stage
constexpr int WR_ROWS = 96;                // staged source rows per stage (32 + tap reach)
constexpr int WR_MAX_STAGES = 4;
constexpr int WR_A_CB = WR_ROWS * 128;     // bytes of one 32-channel block of the source tile (12 KB, 1 KB aligned)

struct WgradReuseArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* gout;
  const float* gout_aux;
  float* partial;          // [splits][Ktot*Cd + Cd]
  int Ls, Ld;              // positions per batch element (source / output)
  int pad_rows, tap_stride, n_taps;
  int kc_cta;              // 32-channel source blocks per CTA (1 or 2)
  int stages_per_batch;    // ceil(Ld / 32)
  int total_stages;        // B * stages_per_batch
  int stages_per_split;
  int n_stages;            // shared-memory ring depth (2..4)
  int want_bias;
};

template <int BN, bool SPLIT, int XFC>
__global__ void __launch_bounds__(UMF_THREADS, 1) conv_wgrad_reuse_kernel(const WgradReuseArgs a) {
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr int NB = BN / 32;
  constexpr int B_PLANE = NB * 4096;
  const int KCC = a.kc_cta;
  const int A_PLANE = KCC * WR_A_CB;
  const int A_BYTES = NP * A_PLANE;
  constexpr int B_BYTES = NP * B_PLANE;
  const int NS = a.n_stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NS * A_BYTES;
  float* s_bias = reinterpret_cast<float*>(sB + NS * B_BYTES);          // [32][BN] column-sum staging
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias + 32 * BN);
  uint64_t* empty_bar = full_bar + WR_MAX_STAGES;
  uint64_t* accum_bar = empty_bar + WR_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.n_taps;
  const int TG = (T + 3) >> 2;                         // tap groups of 4
  const int d = a.tap_stride;
  const int r_in = WR_POS + (4 * TG - 1) * d;          // staged rows (<= WR_ROWS)
  const int cb0 = blockIdx.x * KCC;                    // first 32-channel source block of this CTA
  const int n0 = blockIdx.y * BN;
  const int s_beg = blockIdx.z * a.stages_per_split;
  const int s_end = min(a.total_stages, s_beg + a.stages_per_split);
  const int n_k = s_end - s_beg;                       // >= 1 by construction
  const int groups = KCC * TG;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < groups * BN) tmem_cols <<= 1;
  constexpr int MMA_WARP = UMF_PRODUCERS / 32;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], UMF_PRODUCERS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  float* part = a.partial + (int64_t)blockIdx.z * (Ktot * g.Cd + g.Cd);

  if (warp < MMA_WARP) {
    // ======================= producers: one 16-byte chunk of rows rsub, rsub + 32, rsub + 64 =======================
    const int chunk = tid & 7;
    const int rsub = tid >> 3;                          // 0..31
    constexpr bool NEED_AUX = (XFC == XFC_GENERIC);
    const bool need_aux = NEED_AUX && xf_needs_aux(g.src_xf);
    const bool gneed_aux = xf_needs_aux(g.dst_xf);
    const bool do_bias = a.want_bias && blockIdx.x == 0;
    float4 va[6], ua[6], vb[NB], ub[NB];
    float bsum[NB][4];
#pragma unroll
    for (int j = 0; j < NB; ++j) bsum[j][0] = bsum[j][1] = bsum[j][2] = bsum[j][3] = 0.f;
    int bidx = s_beg / a.stages_per_batch;              // batch element / first position of the next stage to gather
    int l0 = (s_beg - bidx * a.stages_per_batch) * WR_POS;

    auto gather = [&]() {
      const int64_t pix0 = (int64_t)bidx * a.Ls;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int r = rsub + 32 * i;
          const int p = l0 - a.pad_rows + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = v;
          if (c < KCC && r < r_in && (unsigned)p < (unsigned)a.Ls) {
            const int coff = (cb0 + c) * 32 + chunk * 4;
            v = __ldg(reinterpret_cast<const float4*>(a.src + (pix0 + p) * g.ld_src + coff));
            if (NEED_AUX && need_aux)
              u = __ldg(reinterpret_cast<const float4*>(a.src_aux + (pix0 + p) * g.ld_saux + coff));
          }
          va[c * 3 + i] = v;
          if (NEED_AUX) ua[c * 3 + i] = u;
        }
      }
      const int lo = l0 + rsub;
      const int64_t m = (int64_t)bidx * a.Ld + lo;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = v;
        if (lo < a.Ld) {
          v = __ldg(reinterpret_cast<const float4*>(a.gout + m * g.ld_dst + n0 + j * 32 + chunk * 4));
          if (gneed_aux) u = __ldg(reinterpret_cast<const float4*>(a.gout_aux + m * g.ld_daux + n0 + j * 32 + chunk * 4));
        }
        vb[j] = v;
        ub[j] = u;
      }
      l0 += WR_POS;
      if (l0 >= a.stages_per_batch * WR_POS) { l0 = 0; ++bidx; }
    };

    gather();
    int s = 0;
    uint32_t ph = 0;
    for (int ks = 0; ks < n_k; ++ks) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* abase = sA + s * A_BYTES;
      uint8_t* bbase = sB + s * B_BYTES;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int r = rsub + 32 * i;
          if (c < KCC && r < r_in) {
            const float4 x = xf4<XFC>(g, va[c * 3 + i], NEED_AUX ? ua[c * 3 + i] : make_float4(0.f, 0.f, 0.f, 0.f));
            uint8_t* dp = abase + c * WR_A_CB + mn_off(r, chunk);
            if (SPLIT) {
              const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
              *reinterpret_cast<float4*>(dp) = hi;
              *reinterpret_cast<float4*>(dp + A_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
            } else {
              *reinterpret_cast<float4*>(dp) = x;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float4 x = vb[j];
        if (g.dst_xf != MSMC_XF_NONE) {
          x.x = apply_xf(g.dst_xf, g.dst_slope, x.x, ub[j].x); x.y = apply_xf(g.dst_xf, g.dst_slope, x.y, ub[j].y);
          x.z = apply_xf(g.dst_xf, g.dst_slope, x.z, ub[j].z); x.w = apply_xf(g.dst_xf, g.dst_slope, x.w, ub[j].w);
        }
        bsum[j][0] += x.x; bsum[j][1] += x.y; bsum[j][2] += x.z; bsum[j][3] += x.w;
        uint8_t* dp = bbase + j * 4096 + mn_off(rsub, chunk);
        if (SPLIT) {
          const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
          *reinterpret_cast<float4*>(dp) = hi;
          *reinterpret_cast<float4*>(dp + B_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
        } else {
          *reinterpret_cast<float4*>(dp) = x;
        }
      }
      if (ks + 1 < n_k) gather();
      publish_and_arrive_warp(&full_bar[s]);
      if (++s == NS) { s = 0; ph ^= 1u; }
    }

    // ================================= epilogue =================================
    if (do_bias) {
      // fixed-order column sums: thread (rsub, chunk) parks its 4 x NB partials, 32 x BN floats in total
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) s_bias[rsub * BN + j * 32 + chunk * 4 + e] = bsum[j][e];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(UMF_PRODUCERS) : "memory");      // producer warps only
    if (do_bias && tid < BN) {
      float t = 0.f;
      for (int r = 0; r < 32; ++r) t += s_bias[r * BN + tid];
      part[Ktot * g.Cd + n0 + tid] = t;
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // TMEM lane = (tap within its group) * 32 + channel; warps w and w + 4 split the BN columns of every group
    const int j_tap = warp & 3;
    constexpr int CHALF = BN / 2;
    const int cbeg = (warp >> 2) * CHALF;
    for (int gi = 0; gi < groups; ++gi) {
      const int c = gi / TG, tg = gi - c * TG;
      const int tap = tg * 4 + j_tap;
      const uint32_t taddr = tmem_base + ((uint32_t)(j_tap * 32) << 16) + (uint32_t)(gi * BN);
      const int64_t krow = (int64_t)tap * g.Cs + (cb0 + c) * 32 + lane;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + CHALF; c0 += 16) {
        float acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);      // warp-collective: issued for every tap slot, stored for real taps
        if (tap < T) {
          float* out = part + krow * g.Cd + n0 + c0;
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            *reinterpret_cast<float4*>(out + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
        }
      }
    }
    tc_fence_before();
  } else {
    // ================================= MMA issuer =================================
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
    {
      const uint32_t elected = elect_one();     // convergent, predicated issue (see umma_tf32_pred)
      int s = 0;
      uint32_t ph = 0;
      // A: the four M blocks are four taps -> LBO = d rows; B: 32-channel N blocks 4 KB apart
      const uint32_t a_desc0 = ((smem_u32(sA) & 0x3FFFF) >> 4) | ((uint32_t)(d * 8) << 16);
      const uint32_t b_desc0 = desc_lo_mn(smem_u32(sB));
      const uint32_t a_plane = (uint32_t)A_PLANE >> 4;
      for (int ks = 0; ks < n_k; ++ks) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t ad_s = a_desc0 + (uint32_t)s * ((uint32_t)A_BYTES >> 4);
        const uint32_t bd_s = b_desc0 + (uint32_t)s * (B_BYTES >> 4);
        for (int gi = 0; gi < groups; ++gi) {
          const int c = gi / TG, tg = gi - c * TG;
          // group start: channel block c, tap tg*4 -> rows shifted by tg*4*d (8 descriptor units per row)
          const uint32_t ad_g = ad_s + (uint32_t)c * (WR_A_CB >> 4) + (uint32_t)(tg * 4 * d) * 8u;
          const uint32_t td = tmem_base + (uint32_t)(gi * BN);
#pragma unroll
          for (int kg = 0; kg < WR_POS / 8; ++kg) {
            const uint32_t a_hi = ad_g + kg * 64u, b_hi = bd_s + kg * 64u;     // 8 positions = 1 KB = 64 units
            const uint32_t acc = (ks > 0 || kg > 0) ? 1u : 0u;
            if (SPLIT) {
              const uint32_t a_lo = a_hi + a_plane, b_lo = b_hi + (B_PLANE >> 4);
              umma_tf32_pred<DESC_HI_MN>(td, a_lo, b_hi, IDESC, acc, elected);
              umma_tf32_pred<DESC_HI_MN>(td, a_hi, b_lo, IDESC, 1u, elected);
              umma_tf32_pred<DESC_HI_MN>(td, a_hi, b_hi, IDESC, 1u, elected);
            } else {
              umma_tf32_pred<DESC_HI_MN>(td, a_hi, b_hi, IDESC, acc, elected);
            }
          }
        }
        umma_commit_pred(&empty_bar[s], elected);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
      umma_commit_pred(accum_bar, elected);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// GEMM-layout weight [T][Cs][Cd] (cd contiguous)  ->  swizzled tile images [t'][Cs'/32][n_tile][BN x 128 B]
//   role 0 (forward)      : n = cd, k = cs, t' = t
//   role 1 (data gradient): n = cs, k = cd, t' = T-1-t   (stride-1 dgrad == forward conv with reversed taps)
//   role 2 (data gradient, conv-transpose form): n = cs, k = cd, t' = t
__global__ void weight_image_kernel(const float* __restrict__ w, float* __restrict__ img, int T, int Cs, int Cd,
                                    int BN, int role, int split) {
  const int Kdim = role ? Cd : Cs;   // reduction channels of this role
  const int Ndim = role ? Cs : Cd;
  const int KC = (Kdim + 31) / 32;   // a ragged last chunk is zero-padded
  const int NT = (Ndim + BN - 1) / BN;
  const int64_t total = (int64_t)T * KC * NT * BN * 32;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    // decode destination element: [t'][kc][nt][row n][chunk c'][j]   (c' = physical 16-byte slot)
    int64_t r = e;
    const int j = (int)(r & 3); r >>= 2;
    const int cphys = (int)(r & 7); r >>= 3;
    const int nrow = (int)(r % BN); r /= BN;
    const int nt = (int)(r % NT); r /= NT;
    const int kc = (int)(r % KC);
    const int tp = (int)(r / KC);
    const int c = cphys ^ (nrow & 7);                 // logical chunk stored in this slot
    const int k = kc * 32 + c * 4 + j;
    const int n = nt * BN + nrow;
    float val = 0.f;
    if (n < Ndim && k < Kdim) {
      const int t = (role == 1) ? (T - 1 - tp) : tp;
      const int cs = role ? n : k, cd = role ? k : n;
      val = w[((int64_t)t * Cs + cs) * Cd + cd];
    }
    // physical address inside a plane: (nrow/8)*1024 + (nrow%8)*128 + cphys*16 + j*4 bytes == nrow*32 + cphys*4 + j
    // floats; with split the tile is [hi plane][lo plane]
    if (split) {
      const int64_t tile = e / (BN * 32), within = e - tile * (BN * 32);
      const float hi = __uint_as_float(__float_as_uint(val) & 0xFFFFE000u);
      img[tile * (2 * BN * 32) + within] = hi;
      img[tile * (2 * BN * 32) + BN * 32 + within] = val - hi;
    } else {
      img[e] = val;
    }
  }
}

// Multi-tensor form of weight_image_kernel: every operand image a sub-network needs in ONE launch.  Job table
// (device, int64 words, 9 per job): w, img, T, Cs, Cd, BN, role, split, blk0 where blk0 = sum over the preceding jobs
// of ceil(elements / 1024); CTA b serves 1024 destination elements of its job (256 threads x 4).
constexpr int IMG_JOB_WORDS = 9;
__global__ void __launch_bounds__(256) weight_image_multi_kernel(const long long* __restrict__ jobs, int n_jobs) {
  int lo = 0, hi = n_jobs - 1;
  const long long blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[(size_t)mid * IMG_JOB_WORDS + 8] <= blk) lo = mid; else hi = mid - 1;
  }
  const long long* jb = jobs + (size_t)lo * IMG_JOB_WORDS;
  const float* __restrict__ w = reinterpret_cast<const float*>(jb[0]);
  float* __restrict__ img = reinterpret_cast<float*>(jb[1]);
  const int T = (int)jb[2], Cs = (int)jb[3], Cd = (int)jb[4], BN = (int)jb[5], role = (int)jb[6], split = (int)jb[7];
  const int Kdim = role ? Cd : Cs;
  const int Ndim = role ? Cs : Cd;
  const int KC = (Kdim + 31) / 32;
  const int NT = (Ndim + BN - 1) / BN;
  const int64_t total = (int64_t)T * KC * NT * BN * 32;
  const int64_t base = (blk - jb[8]) * 1024;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t e = base + q * 256 + threadIdx.x;
    if (e >= total) break;
    int64_t r = e;
    const int j = (int)(r & 3); r >>= 2;
    const int cphys = (int)(r & 7); r >>= 3;
    const int nrow = (int)(r % BN); r /= BN;
    const int nt = (int)(r % NT); r /= NT;
    const int kc = (int)(r % KC);
    const int tp = (int)(r / KC);
    const int c = cphys ^ (nrow & 7);
    const int k = kc * 32 + c * 4 + j;
    const int n = nt * BN + nrow;
    float val = 0.f;
    if (n < Ndim && k < Kdim) {
      const int t = (role == 1) ? (T - 1 - tp) : tp;
      const int cs = role ? n : k, cd = role ? k : n;
      val = w[((int64_t)t * Cs + cs) * Cd + cd];
    }
    if (split) {
      const int64_t tile = e / (BN * 32), within = e - tile * (BN * 32);
      const float hi_v = __uint_as_float(__float_as_uint(val) & 0xFFFFE000u);
      img[tile * (2 * BN * 32) + within] = hi_v;
      img[tile * (2 * BN * 32) + BN * 32 + within] = val - hi_v;
    } else {
      img[e] = val;
    }
  }
}

}  // namespace

// N tile.  Measured on B200 (profiles/r01_sweep_bn.txt): the tensor pipe retires a 128 x N x 8 TF32 MMA in roughly
// the same time for N = 32 and N = 128, so the widest tile wins almost everywhere -- even when it leaves fewer CTAs
// than SMs (FFN-2 at M = 3840: 60 CTAs of N = 128 take 61 us, 240 CTAs of N = 32 take 82 us).  Only when the grid
// would shrink below ~16 CTAs does the narrower tile's extra parallelism pay (whole step: 41.45 ms at a threshold
// of 32, 41.09 ms at 16, 42.08 ms at 64).
int umma_pick_bn(int cd, int64_t rows) {
  // bring-up / sweep override (profiles/sweep_bn.py): MSMC_FORCE_BN=32|64|128, read per call
  if (const char* e = getenv("MSMC_FORCE_BN")) {
    const int f = atoi(e);
    if ((f == 32 || f == 64 || f == 128) && !(f > 32 && cd <= f / 2)) return f;
  }
  static const int min_ctas = [] { const char* e = getenv("MSMC_BN_MIN_CTAS"); return e ? atoi(e) : 16; }();
  const int64_t mt = ceil_div64(rows, UM_BM);
  const int cands[3] = {128, 64, 32};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 32 && cd <= bn / 2) continue;               // do not pad N by more than 2x
    if (mt * ceil_div(cd, bn) >= min_ctas || bn == 32) return bn;
  }
  return 32;
}

}  // namespace msmc

using namespace msmc;

extern "C" int msmc_umma_tile_n(int32_t out_channels, int64_t rows) { return umma_pick_bn(out_channels, rows); }

extern "C" int64_t msmc_weight_image_elems(int32_t T, int32_t Cs, int32_t Cd, int32_t role, int32_t split,
                                           int32_t BN) {
  const int Kdim = role ? Cd : Cs, Ndim = role ? Cs : Cd;
  if (Kdim % 4 != 0 || (BN != 32 && BN != 64 && BN != 128)) return -1;
  return (int64_t)T * ceil_div(Kdim, 32) * ceil_div(Ndim, BN) * BN * 32 * (split ? 2 : 1);
}

extern "C" int msmc_weight_image(const float* w_gemm, float* image, int32_t T, int32_t Cs, int32_t Cd, int32_t role,
                                 int32_t split, int32_t BN, void* stream) {
  MSMC_REQUIRE(w_gemm && image && T > 0 && Cs > 0 && Cd > 0);
  const int64_t total = msmc_weight_image_elems(T, Cs, Cd, role, 0, BN);
  MSMC_REQUIRE(total > 0);
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 8);
  weight_image_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_gemm, image, T, Cs, Cd, BN, role, split);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_weight_image_multi(const int64_t* jobs, int32_t n_jobs, int64_t total_blocks, void* stream) {
  MSMC_REQUIRE(jobs && n_jobs > 0 && total_blocks > 0 && total_blocks < ((int64_t)1 << 31));
  weight_image_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(jobs), n_jobs);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_conv_forward_umma(const msmc_conv_geom* gp, const float* src, const float* src_aux,
                                      const float* wimg, const float* bias, const float* residual,
                                      const float* dst_aux, float* dst, int32_t split, int32_t BN, void* stream) {
  MSMC_REQUIRE(gp && src && wimg && dst);
  const msmc_conv_geom& g = *gp;
  MSMC_REQUIRE(!(g.transposed && g.pad_reflect));
  MSMC_REQUIRE(g.Cs % 4 == 0 && g.Cs >= UM_MIN_CS && g.ld_src % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0);
  MSMC_REQUIRE(!xf_needs_aux(g.src_xf) ||
               (src_aux && g.ld_saux % 4 == 0 && (reinterpret_cast<uintptr_t>(src_aux) & 15) == 0));
  MSMC_REQUIRE(!xf_needs_aux(g.dst_xf) || dst_aux);
  MSMC_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (g.pad_reflect) MSMC_REQUIRE((g.ph == 0 || g.ph < g.Hs) && (g.pw == 0 || g.pw < g.Ws));
  const int64_t M = g.transposed ? (int64_t)g.B * ceil_div(g.Hd, g.sh) * ceil_div(g.Wd, g.sw)
                                 : (int64_t)g.B * g.Hd * g.Wd;
  const int bn = BN;
  MSMC_REQUIRE(bn == 32 || bn == 64 || bn == 128);
  if (g.transposed) MSMC_REQUIRE(ceil_div(g.KH, g.sh) * ceil_div(g.KW, g.sw) <= UM_MAX_PHASE_TAPS);
  dim3 grid((unsigned)ceil_div64(M, UM_BM), (unsigned)ceil_div(g.Cd, bn),
            (unsigned)(g.transposed ? g.sh * g.sw : 1));
  UmmaArgs a;
  a.g = g; a.src = src; a.src_aux = src_aux; a.wimg = wimg; a.bias = bias; a.residual = residual;
  a.dst_aux = dst_aux; a.dst = dst;
  cudaStream_t st = (cudaStream_t)stream;
  const int xfc = g.src_xf == MSMC_XF_NONE ? XFC_NONE
                  : (g.src_xf == MSMC_XF_LRELU && g.src_slope > 0.f && g.src_slope < 1.f) ? XFC_LRELU : XFC_GENERIC;
#define LAUNCH_UMMA_X(BN_, SPLIT_, ST_, X_)                                                                      \
  do {                                                                                                           \
    const size_t smem = 1024 + (size_t)ST_ * (SPLIT_ ? 2 : 1) * (UM_BM * 128 + BN_ * 128) + (2 * ST_ + 1) * 8 + 16; \
    if (g.transposed) {                                                                                          \
      cudaFuncSetAttribute(conv_umma_kernel<BN_, SPLIT_, ST_, X_, true>,                                         \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                              \
      conv_umma_kernel<BN_, SPLIT_, ST_, X_, true><<<grid, UmmaFwdCfg<BN_, ST_>::PW * 32 + 32, smem, st>>>(a);  \
    } else {                                                                                                     \
      cudaFuncSetAttribute(conv_umma_kernel<BN_, SPLIT_, ST_, X_, false>,                                        \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                              \
      conv_umma_kernel<BN_, SPLIT_, ST_, X_, false><<<grid, UmmaFwdCfg<BN_, ST_>::PW * 32 + 32, smem, st>>>(a); \
    }                                                                                                            \
  } while (0)
#define LAUNCH_UMMA(BN_, SPLIT_, ST_)                                \
  do {                                                               \
    if (xfc == XFC_NONE) LAUNCH_UMMA_X(BN_, SPLIT_, ST_, XFC_NONE);  \
    else if (xfc == XFC_LRELU) LAUNCH_UMMA_X(BN_, SPLIT_, ST_, XFC_LRELU); \
    else LAUNCH_UMMA_X(BN_, SPLIT_, ST_, XFC_GENERIC);               \
  } while (0)
  // short reduction loops (the M-heavy, few-channel layers) are dominated by per-CTA prologue / epilogue latency:
  // a 2-stage ring halves the shared-memory footprint so two CTAs share an SM and overlap those phases
  const int n_k_est = g.KH * g.KW * ceil_div(g.Cs, UM_BK) / (g.transposed ? g.sh * g.sw : 1);
  // measured on the full train step: two co-resident CTAs with 2-stage rings beat one CTA with a 4-stage ring for
  // every BN <= 64 shape (114.5 -> 111.5 ms/step), so "shallow" is the default; MSMC_UMMA_SHALLOW=2 restores the deep ring
  static const int shallow_mode = [] { const char* e = getenv("MSMC_UMMA_SHALLOW"); return e ? atoi(e) : 1; }();
  const bool shallow = shallow_mode == 1 ? true : (shallow_mode == 2 ? false : n_k_est <= 24);
  if (split) {
    switch (bn) {
      case 32: if (shallow) LAUNCH_UMMA(32, true, 2); else LAUNCH_UMMA(32, true, 4); break;   // 2 or 4 x 40 KB
      case 64: if (shallow) LAUNCH_UMMA(64, true, 2); else LAUNCH_UMMA(64, true, 4); break;   // 2 or 4 x 48 KB
      default: LAUNCH_UMMA(128, true, 3); break;                                               // 3 x 64 KB
    }
  } else {
    switch (bn) {
      case 32: if (shallow) LAUNCH_UMMA(32, false, 2); else LAUNCH_UMMA(32, false, 4); break;
      case 64: if (shallow) LAUNCH_UMMA(64, false, 2); else LAUNCH_UMMA(64, false, 4); break;
      default: LAUNCH_UMMA(128, false, 4); break;
    }
  }
#undef LAUNCH_UMMA
#undef LAUNCH_UMMA_X
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

namespace {
int umma_wgrad_bn(int cd) { return cd <= 32 ? 32 : (cd <= 64 ? 64 : 128); }
int umma_wgrad_splits(const msmc_conv_geom& g, int bn) {
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int RB = g.KH * g.KW * ceil_div(g.Cs, 32);
  const int64_t tiles = (int64_t)ceil_div(RB, 4) * ceil_div(g.Cd, bn);
  // the producers are load-latency bound, so the BN <= 64 variants keep two CTAs per SM (2-stage rings); size the
  // position split so that all CTAs are co-resident in one wave (a third, nearly empty wave cost 33 % before)
  const int64_t slots = (int64_t)num_sms() * 2;
  int64_t s = std::max<int64_t>(1, slots / tiles);
  const int64_t max_by_rows = std::max<int64_t>(1, M / 256);      // at least 8 stages per CTA
  if (s > max_by_rows) s = max_by_rows;
  const int64_t per = ((int64_t)g.KH * g.KW * g.Cs * g.Cd + g.Cd) * (int64_t)sizeof(float);
  const int64_t cap = ((int64_t)128 << 20) / per;
  if (s > cap) s = cap;
  return (int)std::max<int64_t>(1, s);
}
}  // namespace

namespace {
bool reuse_eligible(const msmc_conv_geom& g, int* tap_stride, int* n_taps, int* pad_rows);

struct WgradReusePlan {
  int tap_stride, n_taps, pad_rows, kc_cta, bn, splits, stages_per_split, total_stages, stages_per_batch, n_stages;
};
bool wgrad_reuse_plan(const msmc_conv_geom& g, bool split, WgradReusePlan* p) {
  static const int enabled = [] { const char* e = getenv("MSMC_WGRAD_REUSE"); return e ? atoi(e) : 1; }();
  if (!enabled) return false;
  if (!reuse_eligible(g, &p->tap_stride, &p->n_taps, &p->pad_rows)) return false;
  if (g.Cs % 32 != 0 || g.Cd % 32 != 0 || p->n_taps < 3) return false;
  const int TG = (p->n_taps + 3) / 4;
  if (WR_POS + (4 * TG - 1) * p->tap_stride > WR_ROWS) return false;
  // N tile: the tensor pipe spends ~20 cycles per MMA on top of the operand reads (measured: 60 / 68 / 85 cycles for
  // N = 32 / 64 / 128), so the widest tile that fits the accumulator groups into the 512 TMEM columns wins
  static const int wide = [] { const char* e = getenv("MSMC_WGRAD_REUSE_BN128"); return e ? atoi(e) : 1; }();
  p->bn = (g.Cd % 64 == 0) ? 64 : 32;
  p->kc_cta = ((g.Cs / 32) % 2 == 0) ? 2 : 1;
  if (wide && g.Cd % 128 == 0) {
    if (p->kc_cta * TG * 128 <= 512) p->bn = 128;
    else if (TG * 128 <= 512) { p->bn = 128; p->kc_cta = 1; }
  }
  if (p->kc_cta * TG * p->bn > 512) p->kc_cta = 1;
  if (p->kc_cta * TG * p->bn > 512) return false;
  const int Ld = g.Hd * g.Wd;
  p->stages_per_batch = ceil_div(Ld, WR_POS);
  p->total_stages = g.B * p->stages_per_batch;
  const int64_t tiles = (int64_t)(g.Cs / 32 / p->kc_cta) * (g.Cd / p->bn);
  int64_t sp = std::max<int64_t>(1, (int64_t)num_sms() / tiles);
  sp = std::min<int64_t>(sp, std::max<int64_t>(1, p->total_stages / 4));
  const int64_t per = ((int64_t)g.KH * g.KW * g.Cs * g.Cd + g.Cd) * (int64_t)sizeof(float);
  sp = std::max<int64_t>(1, std::min<int64_t>(sp, ((int64_t)128 << 20) / per));
  p->stages_per_split = (int)ceil_div64(p->total_stages, sp);
  p->splits = ceil_div(p->total_stages, p->stages_per_split);
  const int stage_bytes = (split ? 2 : 1) * (p->kc_cta * WR_A_CB + (p->bn / 32) * 4096);
  p->n_stages = std::min(WR_MAX_STAGES, (200 * 1024 - 32 * p->bn * 4) / stage_bytes);
  return p->n_stages >= 2;
}
}  // namespace

extern "C" int64_t msmc_conv_wgrad_umma_workspace(const msmc_conv_geom* gp) {
  if (!gp || gp->Cs % 4 != 0 || gp->Cs < UM_MIN_CS) return -1;
  const msmc_conv_geom& g = *gp;
  {
    WgradReusePlan rp;
    if (wgrad_reuse_plan(g, true, &rp))
      return (int64_t)rp.splits * ((int64_t)g.KH * g.KW * g.Cs * g.Cd + g.Cd) * (int64_t)sizeof(float);
  }
  const int bn = umma_wgrad_bn(g.Cd);
  return (int64_t)umma_wgrad_splits(g, bn) * ((int64_t)g.KH * g.KW * g.Cs * g.Cd + g.Cd) * (int64_t)sizeof(float);
}

extern "C" int msmc_conv_wgrad_umma(const msmc_conv_geom* gp, const float* src, const float* src_aux,
                                    const float* gout, const float* gout_aux, float* dw, float* dbias,
                                    float* workspace, int64_t workspace_bytes, int32_t split, void* stream) {
  MSMC_REQUIRE(gp && src && gout && dw && workspace);
  const msmc_conv_geom& g = *gp;
  MSMC_REQUIRE(!g.transposed && g.Cs % 4 == 0 && g.Cs >= UM_MIN_CS && g.ld_src % 4 == 0 &&
               (reinterpret_cast<uintptr_t>(src) & 15) == 0);
  MSMC_REQUIRE(!xf_needs_aux(g.src_xf) ||
               (src_aux && g.ld_saux % 4 == 0 && (reinterpret_cast<uintptr_t>(src_aux) & 15) == 0));
  MSMC_REQUIRE(!xf_needs_aux(g.dst_xf) || gout_aux);
  const int64_t Ktot = (int64_t)g.KH * g.KW * g.Cs;
  {
    WgradReusePlan rp;
    const bool vec_ok = (g.ld_dst % 4 == 0) && (reinterpret_cast<uintptr_t>(gout) & 15) == 0 &&
                        (!xf_needs_aux(g.dst_xf) ||
                         ((g.ld_daux % 4 == 0) && (reinterpret_cast<uintptr_t>(gout_aux) & 15) == 0));
    // the plan (and the workspace it implies) never depends on `split`: both precisions use the same slicing
    if (vec_ok && wgrad_reuse_plan(g, true, &rp)) {
      MSMC_REQUIRE(workspace_bytes >= (int64_t)rp.splits * (Ktot * g.Cd + g.Cd) * (int64_t)sizeof(float));
      WgradReuseArgs ra;
      ra.g = g; ra.src = src; ra.src_aux = src_aux; ra.gout = gout; ra.gout_aux = gout_aux; ra.partial = workspace;
      ra.Ls = g.Hs * g.Ws; ra.Ld = g.Hd * g.Wd;
      ra.pad_rows = rp.pad_rows; ra.tap_stride = rp.tap_stride; ra.n_taps = rp.n_taps; ra.kc_cta = rp.kc_cta;
      ra.stages_per_batch = rp.stages_per_batch; ra.total_stages = rp.total_stages;
      ra.stages_per_split = rp.stages_per_split; ra.n_stages = rp.n_stages;
      ra.want_bias = dbias != nullptr;
      dim3 rgrid((unsigned)(g.Cs / 32 / rp.kc_cta), (unsigned)(g.Cd / rp.bn), (unsigned)rp.splits);
      const size_t rsmem = 1024 + (size_t)rp.n_stages * (split ? 2 : 1) * (rp.kc_cta * WR_A_CB + (rp.bn / 32) * 4096) +
                           32 * rp.bn * 4 + (2 * WR_MAX_STAGES + 1) * 8 + 16;
      cudaStream_t rst = (cudaStream_t)stream;
      const int rxfc = g.src_xf == MSMC_XF_NONE ? XFC_NONE
                       : (g.src_xf == MSMC_XF_LRELU && g.src_slope > 0.f && g.src_slope < 1.f) ? XFC_LRELU
                                                                                                : XFC_GENERIC;
#define LAUNCH_WR_X(BN_, SPLIT_, X_)                                                                     \
  do {                                                                                                   \
    cudaFuncSetAttribute(conv_wgrad_reuse_kernel<BN_, SPLIT_, X_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)rsmem);                                                                    \
    conv_wgrad_reuse_kernel<BN_, SPLIT_, X_><<<rgrid, UMF_THREADS, rsmem, rst>>>(ra);                    \
  } while (0)
#define LAUNCH_WR(BN_, SPLIT_)                                         \
  do {                                                                 \
    if (rxfc == XFC_NONE) LAUNCH_WR_X(BN_, SPLIT_, XFC_NONE);          \
    else if (rxfc == XFC_LRELU) LAUNCH_WR_X(BN_, SPLIT_, XFC_LRELU);   \
    else LAUNCH_WR_X(BN_, SPLIT_, XFC_GENERIC);                        \
  } while (0)
      if (split) {
        if (rp.bn == 128) LAUNCH_WR(128, true); else if (rp.bn == 64) LAUNCH_WR(64, true); else LAUNCH_WR(32, true);
      } else {
        if (rp.bn == 128) LAUNCH_WR(128, false); else if (rp.bn == 64) LAUNCH_WR(64, false); else LAUNCH_WR(32, false);
      }
#undef LAUNCH_WR
#undef LAUNCH_WR_X
      MSMC_CHECK_LAUNCH();
      return launch_wgrad_reduce(g, workspace, rp.splits, dw, dbias, stream);
    }
  }
  const int bn = umma_wgrad_bn(g.Cd);
  const int splits = umma_wgrad_splits(g, bn);
  MSMC_REQUIRE(workspace_bytes >= (int64_t)splits * (Ktot * g.Cd + g.Cd) * (int64_t)sizeof(float));
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  UmmaWgradArgs a;
  a.g = g; a.src = src; a.src_aux = src_aux; a.gout = gout; a.gout_aux = gout_aux; a.partial = workspace;
  a.rows_per_split = ceil_div64(ceil_div64(M, splits), 32) * 32;
  a.want_bias = dbias != nullptr;
  {
    const char* e = getenv("MSMC_WGRAD_DRY");
    a.dry = e ? atoi(e) : 0;
  }
  MSMC_REQUIRE(M < ((int64_t)1 << 31));
  a.div_hw = make_fastdiv((uint32_t)(g.Hd * g.Wd));
  a.div_w = make_fastdiv((uint32_t)g.Wd);
  a.gvec = (g.ld_dst % 4 == 0) && (reinterpret_cast<uintptr_t>(gout) & 15) == 0 &&
           (!xf_needs_aux(g.dst_xf) || ((g.ld_daux % 4 == 0) && (reinterpret_cast<uintptr_t>(gout_aux) & 15) == 0));
  // every split must own at least one position so that each partial tile is written
  const int eff_splits = (int)ceil_div64(M, a.rows_per_split);
  const int RB = g.KH * g.KW * ceil_div(g.Cs, 32);
  dim3 grid((unsigned)ceil_div(RB, 4), (unsigned)ceil_div(g.Cd, bn), (unsigned)eff_splits);
  cudaStream_t st = (cudaStream_t)stream;
  const int xfc = g.src_xf == MSMC_XF_NONE ? XFC_NONE
                  : (g.src_xf == MSMC_XF_LRELU && g.src_slope > 0.f && g.src_slope < 1.f) ? XFC_LRELU : XFC_GENERIC;
#define LAUNCH_WG_X(BN_, SPLIT_, ST_, X_)                                                                       \
  do {                                                                                                          \
    const size_t smem = 1024 + (size_t)ST_ * (SPLIT_ ? 2 : 1) * (4 * 4096 + (BN_ / 32) * 4096) +                \
                        (2 * ST_ + 1) * 8 + 16;                                                                 \
    cudaFuncSetAttribute(conv_wgrad_umma_kernel<BN_, SPLIT_, ST_, X_>,                                          \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                               \
    conv_wgrad_umma_kernel<BN_, SPLIT_, ST_, X_><<<grid, (BN_ == 128 ? 16 : 8) * 32 + 32, smem, st>>>(a);       \
  } while (0)
#define LAUNCH_WG(BN_, SPLIT_, ST_)                                        \
  do {                                                                     \
    if (xfc == XFC_NONE) LAUNCH_WG_X(BN_, SPLIT_, ST_, XFC_NONE);          \
    else if (xfc == XFC_LRELU) LAUNCH_WG_X(BN_, SPLIT_, ST_, XFC_LRELU);   \
    else LAUNCH_WG_X(BN_, SPLIT_, ST_, XFC_GENERIC);                       \
  } while (0)
  if (split) {
    switch (bn) {
      case 32: LAUNCH_WG(32, true, 2); break;       // 2 x 40 KB, two CTAs per SM
      case 64: LAUNCH_WG(64, true, 2); break;       // 2 x 48 KB, two CTAs per SM
      default: LAUNCH_WG(128, true, 3); break;      // 3 x 64 KB
    }
  } else {
    switch (bn) {
      case 32: LAUNCH_WG(32, false, 4); break;      // 4 x 20 KB
      case 64: LAUNCH_WG(64, false, 4); break;      // 4 x 24 KB
      default: LAUNCH_WG(128, false, 4); break;
    }
  }
#undef LAUNCH_WG
#undef LAUNCH_WG_X
  MSMC_CHECK_LAUNCH();
  return launch_wgrad_reduce(g, workspace, eff_splits, dw, dbias, stream);
}

namespace {
bool reuse_eligible(const msmc_conv_geom& g, int* tap_stride, int* n_taps, int* pad_rows) {
  if (g.transposed || g.pad_reflect || g.sh != 1 || g.sw != 1) return false;
  int ts, nt, pr;
  if (g.Hs == 1 && g.Hd == 1 && g.KH == 1) { ts = g.dw; nt = g.KW; pr = g.pw; }
  else if (g.KW == 1 && g.pw == 0 && g.Ws == g.Wd) { ts = g.dh * g.Ws; nt = g.KH; pr = g.ph * g.Ws; }
  else return false;
  if (nt < 2 || (nt - 1) * ts > RU_ROWS - UM_BM) return false;
  *tap_stride = ts; *n_taps = nt; *pad_rows = pr;
  return true;
}
}  // namespace

extern "C" int msmc_conv_reuse_eligible(const msmc_conv_geom* gp) {
  int a, b, c;
  return gp && gp->Cs % 4 == 0 && gp->Cs >= UM_MIN_CS && reuse_eligible(*gp, &a, &b, &c) ? 1 : 0;
}

extern "C" int msmc_conv_forward_umma_reuse(const msmc_conv_geom* gp, const float* src, const float* src_aux,
                                            const float* wimg, const float* bias, const float* residual,
                                            const float* dst_aux, float* dst, int32_t split, int32_t BN,
                                            void* stream) {
  MSMC_REQUIRE(gp && src && wimg && dst);
  const msmc_conv_geom& g = *gp;
  ReuseArgs a;
  MSMC_REQUIRE(reuse_eligible(g, &a.tap_stride, &a.n_taps, &a.pad_rows));
  MSMC_REQUIRE(g.Cs % 4 == 0 && g.Cs >= UM_MIN_CS && g.ld_src % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0);
  MSMC_REQUIRE(!xf_needs_aux(g.src_xf) ||
               (src_aux && g.ld_saux % 4 == 0 && (reinterpret_cast<uintptr_t>(src_aux) & 15) == 0));
  MSMC_REQUIRE(!xf_needs_aux(g.dst_xf) || dst_aux);
  MSMC_REQUIRE(BN == 32 || BN == 64 || BN == 128);
  {
    // persistent warp-specialised schedule (conv_persist.cu); MSMC_REUSE_PERSIST=0 keeps the one-tile-per-CTA kernel
    // Measured on B200 (profiles/r02_bench_reuse_d.txt): the persistent schedule wins where one launch holds several
    // rounds of long items (many position tiles x >= 7 taps: the 32/64-channel MRF convs, 50 vs 58 us and 64 vs 72 us
    // at k = 11); short items (k = 3) and grids of about one wave are faster one tile per CTA, two CTAs per SM.
    // MSMC_REUSE_PERSIST = 0 / 1 forces one or the other (read per call: the tests switch it).
    const char* e_persist = getenv("MSMC_REUSE_PERSIST");
    const int64_t tiles128 = (int64_t)g.B * ceil_div(g.Hd * g.Wd, UM_BM) * ceil_div(g.Cd, BN);
    const bool want_persist = e_persist ? atoi(e_persist) != 0 : (a.n_taps >= 7 && tiles128 >= 3 * (int64_t)num_sms());
    if (want_persist) {
      const int rc = conv_reuse_persistent(g, src, src_aux, wimg, bias, residual, dst_aux, dst, a.tap_stride,
                                           a.n_taps, a.pad_rows, split, BN, stream);
      if (rc != MSMC_ERR_UNSUPPORTED) return rc;
    }
  }
  a.g = g; a.src = src; a.src_aux = src_aux; a.wimg = wimg; a.bias = bias; a.residual = residual;
  a.dst_aux = dst_aux; a.dst = dst;
  a.Ls = g.Hs * g.Ws; a.Ld = g.Hd * g.Wd;
  a.tiles_per_batch = ceil_div(a.Ld, UM_BM);
  // split-K clusters where a launch has few output tiles and a long channel loop (each CTA then runs KC/KS chunks):
  // as many splits (2 or 4) as keep the grid within one wave and leave >= 4 chunks per CTA.  MSMC_REUSE_KSPLIT=1/2/4
  // forces it (read per call: the A/B tools switch it).
  const int KCh = ceil_div(g.Cs, UM_BK);
  const int64_t out_tiles = (int64_t)a.tiles_per_batch * g.B * ceil_div(g.Cd, BN);
  int KS = 1;
  while (KS < 4 && out_tiles * (KS * 2) <= num_sms() && KCh / (KS * 2) >= 4) KS *= 2;
  if (const char* e_ks = getenv("MSMC_REUSE_KSPLIT")) {
    const int v = atoi(e_ks);
    if (v == 1 || ((v == 2 || v == 4) && KCh >= v)) KS = v;
  }
  a.k_splits = KS;
  dim3 grid((unsigned)(a.tiles_per_batch * g.B), (unsigned)ceil_div(g.Cd, BN), (unsigned)KS);
  cudaStream_t st = (cudaStream_t)stream;
  const int xfc = g.src_xf == MSMC_XF_NONE ? XFC_NONE
                  : (g.src_xf == MSMC_XF_LRELU && g.src_slope > 0.f && g.src_slope < 1.f) ? XFC_LRELU : XFC_GENERIC;
#define LAUNCH_RU_X(BN_, SPLIT_, NA_, NB_, X_)                                                                    \
  do {                                                                                                            \
    const size_t smem = 1024 + (size_t)(SPLIT_ ? 2 : 1) * ((size_t)NA_ * RU_ROWS * 128 + (size_t)NB_ * BN_ * 128) + \
                        (2 * NA_ + 2 * NB_ + 1) * 8 + 16;                                                         \
    /* the partial accumulators of the split-K peers land in the operand stages */                               \
    if ((size_t)(KS - 1) * BN_ * UM_BM * 4 >                                                                      \
        (size_t)(SPLIT_ ? 2 : 1) * ((size_t)NA_ * RU_ROWS * 128 + (size_t)NB_ * BN_ * 128))                       \
      return MSMC_ERR_UNSUPPORTED;                                                                                \
    cudaFuncSetAttribute(conv_umma_reuse_kernel<BN_, SPLIT_, NA_, NB_, X_>,                                       \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                 \
    cudaLaunchConfig_t cfg = {};                                                                                  \
    cfg.gridDim = grid; cfg.blockDim = dim3(ReuseCfg<BN_, NA_>::THREADS, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;      \
    cudaLaunchAttribute attr[1];                                                                                  \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                             \
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)KS;          \
    cfg.attrs = attr; cfg.numAttrs = KS > 1 ? 1 : 0;                                                              \
    if (cudaLaunchKernelEx(&cfg, conv_umma_reuse_kernel<BN_, SPLIT_, NA_, NB_, X_>, a) != cudaSuccess)            \
      return MSMC_ERR_LAUNCH;                                                                                     \
  } while (0)
#define LAUNCH_RU(BN_, SPLIT_, NA_, NB_)                                       \
  do {                                                                         \
    if (xfc == XFC_NONE) LAUNCH_RU_X(BN_, SPLIT_, NA_, NB_, XFC_NONE);         \
    else if (xfc == XFC_LRELU) LAUNCH_RU_X(BN_, SPLIT_, NA_, NB_, XFC_LRELU);  \
    else LAUNCH_RU_X(BN_, SPLIT_, NA_, NB_, XFC_GENERIC);                      \
  } while (0)
  static const int two_ctas = [] { const char* e = getenv("MSMC_REUSE_TWO"); return e ? atoi(e) : 1; }();
  if (split) {
    switch (BN) {
      case 32: if (two_ctas) LAUNCH_RU(32, true, 1, 6); else LAUNCH_RU(32, true, 3, 6); break;   // 48|144 KB + 48 KB
      case 64: if (two_ctas) LAUNCH_RU(64, true, 1, 3); else LAUNCH_RU(64, true, 3, 4); break;   // 48|144 KB + 48|64 KB
      default: {
        // MSMC_REUSE128_TWO=1 (experiment): single operand stage + 2 weight stages = 112 KB -> two CTAs per SM
        static const int two128 = [] { const char* e = getenv("MSMC_REUSE128_TWO"); return e ? atoi(e) : 0; }();
        if (two128) LAUNCH_RU(128, true, 1, 2); else LAUNCH_RU(128, true, 2, 3);    //  96 KB + 96 KB
        break;
      }
    }
  } else {
    switch (BN) {
      case 32: LAUNCH_RU(32, false, 3, 6); break;
      case 64: LAUNCH_RU(64, false, 3, 6); break;
      default: LAUNCH_RU(128, false, 3, 6); break;
    }
  }
#undef LAUNCH_RU
#undef LAUNCH_RU_X
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
