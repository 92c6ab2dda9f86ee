// Persistent, warp-specialised tap-reuse convolution on the 5th-gen tensor cores (tcgen05.mma kind::tf32, 3xTF32).
//
// Same contraction and operand layouts as conv_umma_reuse_kernel (conv_umma.cu): stride-1 convolutions whose taps
// are constant row shifts in flattened position space; one staged operand tile serves every tap through shifted
// shared-memory descriptors; weights arrive as pre-swizzled tile images by cp.async.bulk.  What changes is the
// schedule.  The round-1 kernel ran ONE 128-row tile per CTA with the phases gather -> split -> MMA -> epilogue in
// series and re-streamed every weight tile from L2 for 128 rows of work: ncu showed the tensor pipe busy 29-37 %
// of the time, the MMA warp waiting on weight tiles (three 16-32 KB bulk copies in flight cover ~0.6 us of MMAs, an
// L2 round trip under load is longer) and grids of 0.4-1.7 waves.  Here
//   * the grid is persistent (one CTA per SM, static round-robin over work items) and every role runs its own loop
//     over the items, coupled only by mbarrier rings:
//       warps 0-7   operand producers (gather -> leaky-ReLU / generic transform -> hi/lo split -> swizzled STS),
//                   loads of the next batches are in flight while the current one is stored (NSETS register sets);
//       warps 8-11  epilogue (tcgen05.ld -> bias / activation / residual -> 16-byte stores) of item i while the MMA
//                   warp already works on item i+1: the accumulators are DOUBLE-BUFFERED in TMEM;
//       warp 12     MMA issuer (one elected lane);
//       warp 13     weight-tile loader (cp.async.bulk ring, as deep as shared memory allows);
//   * one work item is NACC x 128 output positions x BN channels: every weight tile fetched from L2 feeds NACC
//     accumulators (NACC x 12 MMAs), which divides the weight traffic and the bulk-copy rate per MMA by NACC;
//   * the operand stage is sized to the layer's real tap reach, so short-reach layers get deeper rings.
#include "umma.cuh"
#include <algorithm>
#include <cstdlib>

namespace msmc {
namespace {

constexpr int PR_PRODUCERS = 256;                 // warps 0-7
constexpr int PR_EPI_WARP0 = 8;                   // warps 8-11 (warp % 4 == TMEM lane quarter)
constexpr int PR_MMA_WARP = 12;
constexpr int PR_THREADS = 14 * 32;
constexpr int PR_RPB = 6;                         // rows per producer thread per batch (batch = 192 rows)
constexpr int PR_MAX_RING = 8;

constexpr int XFC_NONE = 0, XFC_LRELU = 1, XFC_GENERIC = 2;

struct PersistArgs {
  msmc_conv_geom g;
  const float* src;
  const float* src_aux;
  const float* wimg;      // [tap][Cs/32][n_tile][BN rows x 128 B, swizzled] (x2 planes when split)
  const float* bias;
  const float* residual;
  const float* dst_aux;
  float* dst;
  int Ls, Ld;             // positions per batch element (source / destination)
  int pad_rows, tap_stride, n_taps;
  int nacc;               // accumulators (128-row sub-tiles) per work item
  int nbuf;               // TMEM accumulator buffers (2 = epilogue overlaps the next item's MMAs)
  int tiles_per_batch;    // ceil(Ld / (nacc * 128))
  int n_tiles;            // ceil(Cd / BN)
  int n_items;            // B * tiles_per_batch * n_tiles
  int r_in;               // staged rows per chunk = nacc*128 + (n_taps-1)*tap_stride
  int a_plane;            // bytes of one operand plane (multiple of 1024)
  int na, nbs;            // ring depths
  int tg;                 // taps per weight-ring slot (one mbarrier round trip per slot)
  uint32_t tmem_cols;
  int dry;                // bring-up / profiling bit mask: 1 = producers skip loads and stores, 2 = weight loader
                          // skips the copies, 4 = no MMAs are issued, 8 = epilogue skips loads / stores (results invalid)
};

template <int XFC>
__device__ __forceinline__ float4 pxf4(const msmc_conv_geom& g, float4 x, float4 ax) {
  if (XFC == XFC_LRELU) {
    const float sl = g.src_slope;       // slope in (0, 1): leaky_relu(x) == max(x, slope * x)
    x.x = fmaxf(x.x, sl * x.x); x.y = fmaxf(x.y, sl * x.y); x.z = fmaxf(x.z, sl * x.z); x.w = fmaxf(x.w, sl * x.w);
  } else if (XFC == XFC_GENERIC) {
    x.x = apply_xf(g.src_xf, g.src_slope, x.x, ax.x);
    x.y = apply_xf(g.src_xf, g.src_slope, x.y, ax.y);
    x.z = apply_xf(g.src_xf, g.src_slope, x.z, ax.z);
    x.w = apply_xf(g.src_xf, g.src_slope, x.w, ax.w);
  }
  return x;
}

// work item -> (batch element, first output position, channel tile); position tiles vary fastest so that the CTAs
// running side by side stream the same weight tiles (L2 hits)
struct Item {
  int b, l0, n_tile;
};
__device__ __forceinline__ Item decode_item(const PersistArgs& a, int item) {
  Item it;
  const int per_n = a.g.B * a.tiles_per_batch;
  it.n_tile = item / per_n;
  const int r = item - it.n_tile * per_n;
  it.b = r / a.tiles_per_batch;
  it.l0 = (r - it.b * a.tiles_per_batch) * (a.nacc * UM_BM);
  return it;
}

template <int BN, bool SPLIT, int XFC, int NSETS>
__global__ void __launch_bounds__(PR_THREADS, 1) conv_reuse_persist_kernel(const PersistArgs a) {
  const msmc_conv_geom& g = a.g;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NP = SPLIT ? 2 : 1;
  constexpr int B_PLANE = BN * 128;
  constexpr int B_BYTES = NP * B_PLANE;
  const int A_PLANE = a.a_plane;
  const int A_BYTES = NP * A_PLANE;
  const int NA = a.na, NBS = a.nbs;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * A_BYTES;
  const int TG = a.tg;
  const int B_SLOT = TG * B_BYTES;
  uint64_t* fa = reinterpret_cast<uint64_t*>(sB + NBS * B_SLOT);
  uint64_t* ea = fa + PR_MAX_RING;
  uint64_t* fb = ea + PR_MAX_RING;
  uint64_t* eb = fb + PR_MAX_RING;
  uint64_t* tf = eb + PR_MAX_RING;      // [2] accumulator buffer full  (tcgen05.commit)
  uint64_t* te = tf + 2;                // [2] accumulator buffer drained (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(te + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = (g.Cs + UM_BK - 1) / UM_BK;      // ragged last chunk (Cs % 4 == 0) is zero-filled
  const int T = a.n_taps;
  const int NACC = a.nacc;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&fa[i], PR_PRODUCERS / 32); mbar_init(&ea[i], 1); }
    for (int i = 0; i < NBS; ++i) { mbar_init(&fb[i], 1); mbar_init(&eb[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tf[i], 1); mbar_init(&te[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PR_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < PR_EPI_WARP0) {
    // ================================= operand producers =================================
    const int chunk = tid & 7;
    const int rsub = tid >> 3;                 // 0..31; rows rsub + 32*i
    const int r8 = rsub & 7;
    constexpr bool NEED_AUX = (XFC == XFC_GENERIC);
    const bool need_aux = NEED_AUX && xf_needs_aux(g.src_xf);
    const int NRB = (a.r_in + 32 * PR_RPB - 1) / (32 * PR_RPB);      // batches per chunk
    const uint32_t dst_off = (uint32_t)(rsub >> 3) * 1024u + (uint32_t)r8 * 128u + (uint32_t)((chunk ^ r8) << 4);

    // cursor over (item, chunk, batch) in the order the tiles are consumed
    int c_item = blockIdx.x, c_kc = 0, c_rb = 0;         // cursor of the next batch to LOAD
    int s_kc = 0, s_rb = 0;                              // cursor of the next batch to STORE (same order)
    int64_t c_pix0 = 0;
    int c_p0 = 0;
    auto load_cursor_setup = [&]() {
      if (c_item < a.n_items) {
        const Item it = decode_item(a, c_item);
        c_pix0 = (int64_t)it.b * a.Ls;
        c_p0 = it.l0 - a.pad_rows;
      }
    };
    load_cursor_setup();
    float4 v[NSETS][PR_RPB], u[NEED_AUX ? NSETS : 1][PR_RPB];

    auto gather = [&](float4 (&vv)[PR_RPB], float4 (&uu)[PR_RPB]) {
      const int coff = c_kc * UM_BK + chunk * 4;
#pragma unroll
      for (int i = 0; i < PR_RPB; ++i) {
        if (a.dry & 1) { vv[i] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        const int r = c_rb * (32 * PR_RPB) + rsub + 32 * i;
        const int p = c_p0 + r;                             // source position inside this batch element
        if (r < a.r_in && (unsigned)p < (unsigned)a.Ls && coff < g.Cs) {
          vv[i] = __ldg(reinterpret_cast<const float4*>(a.src + (c_pix0 + p) * g.ld_src + coff));
          if (NEED_AUX && need_aux)
            uu[i] = __ldg(reinterpret_cast<const float4*>(a.src_aux + (c_pix0 + p) * g.ld_saux + coff));
        } else {
          vv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (NEED_AUX) uu[i] = vv[i];
        }
      }
      if (++c_rb == NRB) {
        c_rb = 0;
        if (++c_kc == KC) {
          c_kc = 0;
          c_item += gridDim.x;
          load_cursor_setup();
        }
      }
    };

    int sa = 0;
    uint32_t pa = 0;
    auto store = [&](float4 (&vv)[PR_RPB], float4 (&uu)[PR_RPB]) {
      if (s_rb == 0) mbar_wait(&ea[sa], pa ^ 1u);          // the MMAs that read this slot have retired
      uint8_t* dstbase = sA + sa * A_BYTES + dst_off + (uint32_t)s_rb * (uint32_t)(32 * PR_RPB * 128);
#pragma unroll
      for (int i = 0; i < PR_RPB; ++i) {
        if (s_rb * (32 * PR_RPB) + rsub + 32 * i < a.r_in && !(a.dry & 1)) {
          const float4 x = pxf4<XFC>(g, vv[i], NEED_AUX ? uu[i] : make_float4(0.f, 0.f, 0.f, 0.f));
          uint8_t* d = dstbase + i * 4096;        // rows advance by 32 -> four 1 KB swizzle atoms
          if (SPLIT) {
            const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
            *reinterpret_cast<float4*>(d) = hi;
            *reinterpret_cast<float4*>(d + A_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
          } else {
            *reinterpret_cast<float4*>(d) = x;
          }
        }
      }
      if (++s_rb == NRB) {
        s_rb = 0;
        publish_and_arrive_warp(&fa[sa]);
        if (++sa == NA) { sa = 0; pa ^= 1u; }
        ++s_kc;
      }
    };

    // total batches this CTA produces
    const int my_items = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_items * KC * NRB;
    // prologue: NSETS - 1 batches in flight
#pragma unroll
    for (int s = 0; s < NSETS - 1; ++s)
      if (s < total) gather(v[s], u[NEED_AUX ? s : 0]);
    for (int q = 0; q < total; q += NSETS) {
#pragma unroll
      for (int s = 0; s < NSETS; ++s) {
        if (q + s < total) {
          const int ls = (s + NSETS - 1) % NSETS;            // register set freed by the previous store
          if (q + s + NSETS - 1 < total) gather(v[ls], u[NEED_AUX ? ls : 0]);
          store(v[s], u[NEED_AUX ? s : 0]);
        }
      }
    }
    (void)s_kc;
  } else if (warp < PR_MMA_WARP) {
    // ================================= epilogue =================================
    const int lane_grp = warp & 3;
    const bool dneed_aux = xf_needs_aux(g.dst_xf);
    const bool vec_ok = (!dneed_aux || ((g.ld_daux & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dst_aux) & 15) == 0)) &&
                        (g.ld_dst & 3) == 0 && (!a.residual || ((g.ld_res & 3) == 0 &&
                                                               (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0)) &&
                        (!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
    uint32_t ph_tf[2] = {0u, 0u};
    int k = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
      const int buf = (a.nbuf == 2) ? (k & 1) : 0;
      const Item it = decode_item(a, item);
      mbar_wait(&tf[buf], ph_tf[buf]);
      ph_tf[buf] ^= 1u;
      tc_fence_after();
      const int n0 = it.n_tile * BN;
      for (int acc_i = 0; acc_i < NACC; ++acc_i) {
        const int l = it.l0 + acc_i * UM_BM + lane_grp * 32 + lane;
        if (it.l0 + acc_i * UM_BM >= a.Ld) break;           // (uniform) sub-tile beyond the sequence: no MMAs ran
        const bool row_ok = l < a.Ld;
        const int64_t m = (int64_t)it.b * a.Ld + l;
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)((buf * NACC + acc_i) * BN);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float acc[16];
          tmem_ld16(taddr + (uint32_t)c0, acc);
          if (!row_ok || (a.dry & 8)) continue;
          const bool full16 = n0 + c0 + 16 <= g.Cd;
          if (full16 && vec_ok) {
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
              a.dst[m * g.ld_dst + n] = x;
            }
          }
        }
      }
      // this warp's TMEM reads of the buffer are complete (tcgen05.wait::ld inside tmem_ld16): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&te[buf]);
    }
  } else if (warp == PR_MMA_WARP) {
    // ================================= MMA issuer =================================
    // The WHOLE warp walks the loop nest with identical (warp-uniform) values and one lane issues.  Every quantity
    // that reaches a tcgen05.mma operand is derived from kernel parameters, loop counters and values laundered
    // through __shfl_sync(.., 0): ptxas then keeps descriptors / TMEM addresses in uniform registers.  (With the loop
    // inside `if (lane == 0)` and the item decode's integer divisions feeding the loop bounds, every MMA was wrapped
    // in a 5 x R2UR.BROADCAST waterfall loop and the issuing thread, not the tensor pipe, set the pace at N <= 64.)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(UM_BM >> 4) << 24);
    {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t ph_te = 0;                                            // bit b = parity of accumulator buffer b
      const uint32_t a_desc0 = desc_lo_k(smem_u32(sA)), b_desc0 = desc_lo_k(smem_u32(sB));
      const uint32_t a_step = (uint32_t)a.tap_stride * 8u;          // rows of 128 B = 8 descriptor units
      const uint32_t a_plane_u = (uint32_t)A_PLANE >> 4;
      const uint32_t a_stage_u = (uint32_t)A_BYTES >> 4;
      const uint32_t elected = elect_one();                         // the same lane issues every MMA and commit
      int k = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
        const int buf = (a.nbuf == 2) ? (k & 1) : 0;
        const Item it = decode_item(a, item);
        int n_acc = (a.Ld - it.l0 + UM_BM - 1) / UM_BM;              // sub-tiles that hold real positions
        if (n_acc > NACC) n_acc = NACC;
        n_acc = __shfl_sync(0xffffffffu, n_acc, 0);                  // provably warp-uniform loop bound
        // the epilogue has drained this buffer (first use of each buffer passes immediately)
        mbar_wait(&te[buf], ((ph_te >> buf) & 1u) ^ 1u);
        ph_te ^= 1u << buf;
        tc_fence_after();
        const uint32_t td0 = tmem_base + (uint32_t)(buf * NACC * BN);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&fa[sa], pa);
          const uint32_t ad_s = a_desc0 + (uint32_t)sa * a_stage_u;
          uint32_t ad_t = ad_s;
          for (int t0 = 0; t0 < T; t0 += TG) {
            const int g_taps = min(TG, T - t0);
            mbar_wait(&fb[sb], pb);
            tc_fence_after();
            uint32_t bd = b_desc0 + (uint32_t)sb * ((uint32_t)B_SLOT >> 4);
            for (int tt = 0; tt < g_taps; ++tt, ad_t += a_step, bd += (B_BYTES >> 4)) {
              const uint32_t first = (kc > 0 || t0 + tt > 0) ? 1u : 0u;
              // convergent, predicated issue (see umma_tf32_pred): all lanes run the same straight-line code
#pragma unroll
              for (int acc_i = 0; acc_i < 4; ++acc_i) {
                if (acc_i < n_acc && !(a.dry & 4)) {
                  const uint32_t ad = ad_t + (uint32_t)acc_i * (UM_BM * 8u), td = td0 + (uint32_t)(acc_i * BN);
#pragma unroll
                  for (int ks = 0; ks < UM_BK / 8; ++ks) {
                    const uint32_t a_hi = ad + 2 * ks, b_hi = bd + 2 * ks;
                    const uint32_t accf = ks > 0 ? 1u : first;
                    if (SPLIT) {
                      const uint32_t a_lo = a_hi + a_plane_u, b_lo = b_hi + (B_PLANE >> 4);
                      umma_tf32_pred<DESC_HI_K>(td, a_lo, b_hi, IDESC, accf, elected);
                      umma_tf32_pred<DESC_HI_K>(td, a_hi, b_lo, IDESC, 1u, elected);
                      umma_tf32_pred<DESC_HI_K>(td, a_hi, b_hi, IDESC, 1u, elected);
                    } else {
                      umma_tf32_pred<DESC_HI_K>(td, a_hi, b_hi, IDESC, accf, elected);
                    }
                  }
                }
              }
            }
            umma_commit_pred(&eb[sb], elected);
            if (t0 + TG >= T) umma_commit_pred(&ea[sa], elected);
            if (t0 + TG >= T && kc == KC - 1) umma_commit_pred(&tf[buf], elected);
            __syncwarp();
            if (++sb == NBS) { sb = 0; pb ^= 1u; }
          }
          if (++sa == NA) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else {
    // ================================= weight-tile loader =================================
    if (lane == 0) {
      int sb = 0;
      uint32_t pb = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const Item it = decode_item(a, item);
        for (int kc = 0; kc < KC; ++kc)
          for (int t0 = 0; t0 < T; t0 += TG) {
            const int g_taps = min(TG, T - t0);
            mbar_wait(&eb[sb], pb ^ 1u);
            if (a.dry & 2) { mbar_arrive(&fb[sb]); if (++sb == NBS) { sb = 0; pb ^= 1u; } continue; }
            mbar_arrive_expect_tx(&fb[sb], (uint32_t)(g_taps * B_BYTES));
            for (int tt = 0; tt < g_taps; ++tt) {
              const float* wsrc = a.wimg + (((int64_t)(t0 + tt) * KC + kc) * a.n_tiles + it.n_tile) * (B_BYTES / 4);
              bulk_g2s(sB + sb * B_SLOT + tt * B_BYTES, wsrc, B_BYTES, &fb[sb]);
            }
            if (++sb == NBS) { sb = 0; pb ^= 1u; }
          }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PR_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

}  // namespace

// Launch plan + launch.  Returns MSMC_ERR_UNSUPPORTED when the shape does not fit (the caller falls back to the
// one-tile-per-CTA kernel).
int conv_reuse_persistent(const msmc_conv_geom& g, const float* src, const float* src_aux, const float* wimg,
                          const float* bias, const float* residual, const float* dst_aux, float* dst,
                          int tap_stride, int n_taps, int pad_rows, int split, int BN, void* stream) {
  PersistArgs a;
  a.g = g; a.src = src; a.src_aux = src_aux; a.wimg = wimg; a.bias = bias; a.residual = residual;
  a.dst_aux = dst_aux; a.dst = dst;
  a.Ls = g.Hs * g.Ws; a.Ld = g.Hd * g.Wd;
  a.pad_rows = pad_rows; a.tap_stride = tap_stride; a.n_taps = n_taps;
  a.n_tiles = ceil_div(g.Cd, BN);
  const int reach = (n_taps - 1) * tap_stride;
  const int np = split ? 2 : 1;
  const int b_stage = np * BN * 128;
  const int budget = 224 * 1024 - 1024 - 512;       // dynamic shared memory minus alignment slack and barriers
  // bring-up / test overrides, read per call: MSMC_PERSIST_NACC = 1|2|4, MSMC_PERSIST_NBS = weight-ring depth cap
  const char* e_nacc = getenv("MSMC_PERSIST_NACC");
  const char* e_nbs = getenv("MSMC_PERSIST_NBS");
  const int force_nacc = e_nacc ? atoi(e_nacc) : 0, force_nbs = e_nbs ? atoi(e_nbs) : 0;
  // Weight-ring slot = up to `tg` taps (<= 32 KB): one mbarrier round trip serves tg x 12 MMAs.  Measured on B200
  // (profiles/r02_bench_reuse_*.txt): the barrier skeleton of this kernel alone costs 12-25 us per launch at one tap
  // per slot, comparable to the MMA time of the 3-tap layers.
  const char* e_tg = getenv("MSMC_PERSIST_TG");
  int tg = std::max(1, std::min(n_taps, (32 * 1024) / b_stage));
  if (e_tg) tg = std::max(1, std::min(n_taps, atoi(e_tg)));
  const int b_slot = tg * b_stage;
  // Accumulators per item.  NACC > 1 re-uses every weight tile for NACC x 12 MMAs; measured, it never pays on this
  // step's shapes (the MMAs, not the weight stream, set the pace and fewer / larger items quantise worse), so the
  // default is 1 and MSMC_PERSIST_NACC keeps the others testable.
  int best_nacc = 0, best_na = 0, best_nbs = 0;
  double best_cost = 0.0;
  for (int nacc = 1; nacc <= 4; nacc *= 2) {
    if (force_nacc ? nacc != force_nacc : nacc != 1) continue;
    if (nacc * BN > 512) continue;
    const int r_in = nacc * UM_BM + reach;
    const int a_stage = np * ceil_div(r_in, 8) * 1024;
    int na = 2;
    if (2 * a_stage + 2 * b_slot > budget) na = 1;
    if (na * a_stage + 2 * b_slot > budget) continue;
    int nbs = std::min(PR_MAX_RING, (budget - na * a_stage) / b_slot);
    if (force_nbs) nbs = std::min(nbs, std::max(2, force_nbs));
    best_nacc = nacc; best_na = na; best_nbs = nbs; best_cost = 0.0;
  }
  (void)best_cost;
  if (best_nacc == 0) return MSMC_ERR_UNSUPPORTED;
  a.nacc = best_nacc; a.na = best_na; a.nbs = best_nbs; a.tg = tg;
  a.r_in = a.nacc * UM_BM + reach;
  a.a_plane = ceil_div(a.r_in, 8) * 1024;
  a.tiles_per_batch = ceil_div(a.Ld, a.nacc * UM_BM);
  a.n_items = g.B * a.tiles_per_batch * a.n_tiles;
  a.nbuf = (2 * a.nacc * BN <= 512) ? 2 : 1;
  uint32_t cols = 32;
  while ((int)cols < a.nbuf * a.nacc * BN) cols <<= 1;
  a.tmem_cols = cols;
  { const char* e = getenv("MSMC_PERSIST_DRY"); a.dry = e ? atoi(e) : 0; }
  // one CTA per SM by construction (TMEM is not shared between co-resident CTAs of this kernel): ask for more than
  // half of the shared memory even when the rings are small
  size_t smem = 1024 + (size_t)a.na * np * a.a_plane + (size_t)a.nbs * b_slot + (4 * PR_MAX_RING + 4) * 8 + 16;
  smem = std::max<size_t>(smem, 118 * 1024);
  const int grid = std::min(a.n_items, num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  const int xfc = g.src_xf == MSMC_XF_NONE ? XFC_NONE
                  : (g.src_xf == MSMC_XF_LRELU && g.src_slope > 0.f && g.src_slope < 1.f) ? XFC_LRELU : XFC_GENERIC;
#define LAUNCH_P_X(BN_, SPLIT_, X_, NS_)                                                                      \
  do {                                                                                                        \
    cudaFuncSetAttribute(conv_reuse_persist_kernel<BN_, SPLIT_, X_, NS_>,                                     \
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                             \
    conv_reuse_persist_kernel<BN_, SPLIT_, X_, NS_><<<grid, PR_THREADS, smem, st>>>(a);                       \
  } while (0)
#define LAUNCH_P(BN_, SPLIT_)                                          \
  do {                                                                 \
    if (xfc == XFC_NONE) LAUNCH_P_X(BN_, SPLIT_, XFC_NONE, 3);         \
    else if (xfc == XFC_LRELU) LAUNCH_P_X(BN_, SPLIT_, XFC_LRELU, 3);  \
    else LAUNCH_P_X(BN_, SPLIT_, XFC_GENERIC, 2);                      \
  } while (0)
  if (split) {
    switch (BN) {
      case 32: LAUNCH_P(32, true); break;
      case 64: LAUNCH_P(64, true); break;
      default: LAUNCH_P(128, true); break;
    }
  } else {
    switch (BN) {
      case 32: LAUNCH_P(32, false); break;
      case 64: LAUNCH_P(64, false); break;
      default: LAUNCH_P(128, false); break;
    }
  }
#undef LAUNCH_P
#undef LAUNCH_P_X
  if (cudaGetLastError() != cudaSuccess) return MSMC_ERR_LAUNCH;
  return MSMC_OK;
}

}  // namespace msmc
