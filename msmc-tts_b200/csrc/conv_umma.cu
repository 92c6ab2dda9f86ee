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
#include "common.cuh"
#include <algorithm>

namespace msmc {
namespace {

constexpr int UM_BM = 128;
constexpr int UM_BK = 32;       // tf32 per stage row (128 B)
constexpr int UM_THREADS = 160;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 = 1 | [32,46) SBO>>4 = 64 (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

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

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int BN, bool SPLIT, int STAGES>
__global__ void __launch_bounds__(UM_THREADS, 1) conv_umma_kernel(const UmmaArgs a) {
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: A stages (16 KB each), B stages (BN*128 B each), barriers, tmem slot
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
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int64_t m0 = (int64_t)blockIdx.x * UM_BM;
  const int n_tile = blockIdx.y;
  const int n_tiles = gridDim.y;
  const int KC = g.Cs / UM_BK;
  const int T = g.KH * g.KW;
  const int n_k = T * KC;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 128 + 1);   // 128 producer arrivals + 1 arrive.expect_tx for the weight tile
      mbar_init(&empty_bar[s], 1);        // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    // allocate BN TMEM columns (power of two >= 32); the address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================================= A producers =================================
    const int64_t m = m0 + tid;
    const bool row_ok = m < M;
    int b = 0, hd = 0, wd = 0;
    if (row_ok) {
      b = (int)(m / ((int64_t)g.Hd * g.Wd));
      const int rem = (int)(m % ((int64_t)g.Hd * g.Wd));
      hd = rem / g.Wd;
      wd = rem - hd * g.Wd;
    }
    const bool need_aux = xf_needs_aux(g.src_xf);
    const int r8 = tid & 7;
    const uint32_t row_off = (uint32_t)(tid >> 3) * 1024u + (uint32_t)r8 * 128u;
    float4 v[8], u[8];

    // raw global loads only (no dependent math), so they stay in flight across the barrier round-trip
    auto gather = [&](int ks) {
      const int t = ks / KC, kc = ks - t * KC;
      const int kh = t / g.KW, kw = t - kh * g.KW;
      int hs = hd * g.sh + kh * g.dh - g.ph;
      int ws = wd * g.sw + kw * g.dw - g.pw;
      bool ok = row_ok;
      if (g.pad_reflect) {
        hs = reflect1(hs, g.Hs);
        ws = reflect1(ws, g.Ws);
      } else {
        ok = ok && hs >= 0 && hs < g.Hs && ws >= 0 && ws < g.Ws;
      }
      if (ok) {
        const int64_t pix = ((int64_t)b * g.Hs + hs) * g.Ws + ws;
        const float4* p = reinterpret_cast<const float4*>(a.src + pix * g.ld_src + kc * UM_BK);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldg(p + c);
        if (need_aux) {
          const float4* q = reinterpret_cast<const float4*>(a.src_aux + pix * g.ld_saux + kc * UM_BK);
#pragma unroll
          for (int c = 0; c < 8; ++c) u[c] = __ldg(q + c);
        }
      } else {
        // padding: every operand transform maps 0 -> 0, so zeros pass through the store path unchanged
#pragma unroll
        for (int c = 0; c < 8; ++c) { v[c] = make_float4(0.f, 0.f, 0.f, 0.f); u[c] = v[c]; }
      }
    };

    gather(0);
    for (int ks = 0; ks < n_k; ++ks) {
      const int s = ks % STAGES;
      const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar[s], B_BYTES);
        const float* wsrc = a.wimg + ((int64_t)ks * n_tiles + n_tile) * (B_BYTES / 4);
        bulk_g2s(sB + s * B_BYTES, wsrc, B_BYTES, &full_bar[s]);
      }
      uint8_t* dstrow = sA + s * A_BYTES + row_off;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 x = v[c];
        if (g.src_xf != MSMC_XF_NONE) {
          const float4 ax = need_aux ? u[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          x.x = apply_xf(g.src_xf, g.src_slope, x.x, ax.x);
          x.y = apply_xf(g.src_xf, g.src_slope, x.y, ax.y);
          x.z = apply_xf(g.src_xf, g.src_slope, x.z, ax.z);
          x.w = apply_xf(g.src_xf, g.src_slope, x.w, ax.w);
        }
        if (SPLIT) {
          const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
          *reinterpret_cast<float4*>(dstrow + ((c ^ r8) << 4)) = hi;
          *reinterpret_cast<float4*>(dstrow + A_PLANE + ((c ^ r8) << 4)) =
              make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
        } else {
          *reinterpret_cast<float4*>(dstrow + ((c ^ r8) << 4)) = x;
        }
      }
      if (ks + 1 < n_k) gather(ks + 1);   // issue the next stage's loads before signalling this one
      fence_proxy_async();
      mbar_arrive(&full_bar[s]);
    }

    // ================================= epilogue =================================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int n0 = n_tile * BN;
    const bool dneed_aux = xf_needs_aux(g.dst_xf);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      if (row_ok) {
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
        if (n0 + c0 + 16 <= g.Cd && (g.ld_dst & 3) == 0) {
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
  } else {
    // ================================= MMA issuer (warp 4) =================================
    // instruction descriptor: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(UM_BM >> 4) << 24);
    if ((tid & 31) == 0) {
      for (int ks = 0; ks < n_k; ++ks) {
        const int s = ks % STAGES;
        const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * A_BYTES);
        const uint32_t b_addr = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < UM_BK / 8; ++k) {
          // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row
          const uint64_t a_hi = make_desc(a_addr + k * 32), b_hi = make_desc(b_addr + k * 32);
          if (SPLIT) {
            const uint64_t a_lo = make_desc(a_addr + A_PLANE + k * 32), b_lo = make_desc(b_addr + B_PLANE + k * 32);
            umma_tf32(tmem_base, a_lo, b_hi, IDESC, (ks > 0 || k > 0) ? 1u : 0u);   // small terms first
            umma_tf32(tmem_base, a_hi, b_lo, IDESC, 1u);
            umma_tf32(tmem_base, a_hi, b_hi, IDESC, 1u);
          } else {
            umma_tf32(tmem_base, a_hi, b_hi, IDESC, (ks > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);   // frees the stage when these MMAs have read it
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}

// GEMM-layout weight [T][Cs][Cd] (cd contiguous)  ->  swizzled tile images [t'][Cs'/32][n_tile][BN x 128 B]
//   role 0 (forward)      : n = cd, k = cs, t' = t
//   role 1 (data gradient): n = cs, k = cd, t' = T-1-t   (stride-1 dgrad == forward conv with reversed taps)
__global__ void weight_image_kernel(const float* __restrict__ w, float* __restrict__ img, int T, int Cs, int Cd,
                                    int BN, int role, int split) {
  const int Kdim = role ? Cd : Cs;   // reduction channels of this role
  const int Ndim = role ? Cs : Cd;
  const int KC = Kdim / 32;
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
    if (n < Ndim) {
      const int t = role ? (T - 1 - tp) : tp;
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

}  // namespace

int umma_pick_bn(int cd) { return cd <= 32 ? 32 : (cd <= 64 ? 64 : 128); }

}  // namespace msmc

using namespace msmc;

extern "C" int msmc_umma_tile_n(int32_t out_channels) { return umma_pick_bn(out_channels); }

extern "C" int64_t msmc_weight_image_elems(int32_t T, int32_t Cs, int32_t Cd, int32_t role, int32_t split) {
  const int Kdim = role ? Cd : Cs, Ndim = role ? Cs : Cd;
  if (Kdim % 32 != 0) return -1;
  const int BN = umma_pick_bn(Ndim);
  return (int64_t)T * (Kdim / 32) * ceil_div(Ndim, BN) * BN * 32 * (split ? 2 : 1);
}

extern "C" int msmc_weight_image(const float* w_gemm, float* image, int32_t T, int32_t Cs, int32_t Cd, int32_t role,
                                 int32_t split, void* stream) {
  MSMC_REQUIRE(w_gemm && image && T > 0 && Cs > 0 && Cd > 0);
  const int64_t total = msmc_weight_image_elems(T, Cs, Cd, role, 0);
  MSMC_REQUIRE(total > 0);
  const int Ndim = role ? Cs : Cd;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 8);
  weight_image_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_gemm, image, T, Cs, Cd, umma_pick_bn(Ndim), role,
                                                                split);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_conv_forward_umma(const msmc_conv_geom* gp, const float* src, const float* src_aux,
                                      const float* wimg, const float* bias, const float* residual,
                                      const float* dst_aux, float* dst, int32_t split, void* stream) {
  MSMC_REQUIRE(gp && src && wimg && dst);
  const msmc_conv_geom& g = *gp;
  MSMC_REQUIRE(!g.transposed);
  MSMC_REQUIRE(g.Cs % UM_BK == 0 && g.ld_src % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0);
  MSMC_REQUIRE(!xf_needs_aux(g.src_xf) ||
               (src_aux && g.ld_saux % 4 == 0 && (reinterpret_cast<uintptr_t>(src_aux) & 15) == 0));
  MSMC_REQUIRE(!xf_needs_aux(g.dst_xf) || dst_aux);
  MSMC_REQUIRE((reinterpret_cast<uintptr_t>(wimg) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (g.pad_reflect) MSMC_REQUIRE((g.ph == 0 || g.ph < g.Hs) && (g.pw == 0 || g.pw < g.Ws));
  const int64_t M = (int64_t)g.B * g.Hd * g.Wd;
  const int bn = umma_pick_bn(g.Cd);
  dim3 grid((unsigned)ceil_div64(M, UM_BM), (unsigned)ceil_div(g.Cd, bn));
  UmmaArgs a;
  a.g = g; a.src = src; a.src_aux = src_aux; a.wimg = wimg; a.bias = bias; a.residual = residual;
  a.dst_aux = dst_aux; a.dst = dst;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_UMMA(BN_, SPLIT_, ST_)                                                                            \
  do {                                                                                                           \
    const size_t smem = 1024 + (size_t)ST_ * (SPLIT_ ? 2 : 1) * (UM_BM * 128 + BN_ * 128) + (2 * ST_ + 1) * 8 + 16; \
    cudaFuncSetAttribute(conv_umma_kernel<BN_, SPLIT_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                         (int)smem);                                                                             \
    conv_umma_kernel<BN_, SPLIT_, ST_><<<grid, UM_THREADS, smem, st>>>(a);                                       \
  } while (0)
  if (split) {
    switch (bn) {
      case 32: LAUNCH_UMMA(32, true, 4); break;    // 4 x 40 KB
      case 64: LAUNCH_UMMA(64, true, 4); break;    // 4 x 48 KB
      default: LAUNCH_UMMA(128, true, 3); break;   // 3 x 64 KB
    }
  } else {
    switch (bn) {
      case 32: LAUNCH_UMMA(32, false, 4); break;
      case 64: LAUNCH_UMMA(64, false, 4); break;
      default: LAUNCH_UMMA(128, false, 4); break;
    }
  }
#undef LAUNCH_UMMA
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
