// Two-phase multi-head VQ search on the tensor cores (msmc_vq_search_umma; validated bit-exact against
// oracle/vq_oracle.c by tests/test_vq_umma_gpu.py).
//
// The CUDA-core search kernels (vq.cu) are fp32-FMA bound: at K = 256 a row costs 65 536 FMAs per 3 360 bytes, so the
// exhaustive search tops out near 28 % of the HBM roofline.  Here
//   phase 1  scores a 128-row tile of one head against all K codewords with tcgen05.mma (3xTF32: fp32-accurate to
//            ~2^-21 relative, accumulators in TMEM): A = z rows, B = the head's codebook, both K-major SWIZZLE_128B
//            (a 128-byte shared-memory row = 32 dims of one row / one codeword), two K halves of 32 dims;
//   phase 2  reads the 128 x K dot products from TMEM, forms dist = (|z|^2 - 2 z.e_k) + |e_k|^2 with the exact
//            (sequential-fma) codeword norms, and checks that the runner-up lies more than 2*delta above the minimum,
//            delta = 2^-20 (|z|^2 + max|e|^2 + 2 |z| max|e|) (tests/test_vq_two_phase_margin.py: then the minimum is
//            the exhaustive search's argmin; true for > 99.9 % of the rows).  Any other row re-scores every codeword
//            within 2*delta of its minimum with the oracle's exact arithmetic and tie rule (lowest index).
// The result is identical to the exhaustive fp32 search by construction.
// PERSISTENT + WARP-SPECIALISED: the heads of a row tile form a thread-block cluster (CTA = head) and a cluster walks
// the row tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  Each CTA stages ITS head's codebook (hi / lo planes,
// exact norms) in shared memory once.  See the kernel for the pipeline.
#include "umma.cuh"
#include <algorithm>
#include <cooperative_groups.h>
#include <cstdlib>

namespace msmc {
namespace {

constexpr int VU_ROWS = 128;                 // rows per CTA (TMEM lanes)
constexpr int VU_DIM = 64;                   // dims per head (two K halves of 32)
constexpr int VU_EPI = 256;                  // warps 0-3 / 4-7: two epilogue groups (even / odd tiles of the CTA)
constexpr int VU_STAGERS = 128;              // warps 8-9: operand staging, warps 10-11: head sum
constexpr int VU_THREADS = VU_EPI + VU_STAGERS + 32;  // + the MMA warp

// ---- cluster / distributed-shared-memory helpers (the heads of a row tile exchange their commitment terms) ----
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
// fire-and-forget remote store that reports its 4 bytes to an mbarrier of the destination CTA (complete_tx): the
// pusher needs no release fence (which would also wait for its global stores) and no arrival of its own
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr),
               "r"(__float_as_uint(v)), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Wait for a barrier whose guarded data was written by OTHER CTAs of the cluster.  The probes are the plain
// (CTA-scope) try_wait -- a cluster-scope acquire in the spin loop compiles to one CCTL.IVALL per probe, which kept
// the L1 empty for every other warp of the SM, and a fence.acq_rel.cluster after it to a MEMBAR.ALL.GPU that also
// drains this warp's own global stores -- followed by ONE cluster-scope acquiring test of the completed phase.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  mbar_wait(bar, parity);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// tcgen05.ld without the wait, and a wait that names the destination registers (so no consumer of them can be
// scheduled above it): lets the scan keep one TMEM load in flight under the compares of the previous chunk
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// one chunk of 16 approximate distances folded into FOUR independent (best, runner-up, index) accumulators (column
// j goes to accumulator j & 3): with one epilogue warp per scheduler a single accumulator is a dependent chain of
// ~30 cycles per column; four chains keep the scan issue-bound (~7 instructions per column).  Branch-free:
//   second = min(second, max(best, d));  index = d < best ? c : index;  best = min(best, d)
__device__ __forceinline__ void scan16(const uint32_t (&acc)[16], const float* __restrict__ ee, int c0, float zz,
                                       float (&best)[4], float (&second)[4], int (&best_c)[4]) {
#pragma unroll
  for (int j4 = 0; j4 < 16; j4 += 4) {
    const float4 e4 = *reinterpret_cast<const float4*>(ee + c0 + j4);     // (same address in every lane: broadcast)
    const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dj = (zz - 2.f * __uint_as_float(acc[j4 + j])) + ev[j];
      second[j] = fminf(second[j], fmaxf(best[j], dj));
      best_c[j] = dj < best[j] ? c0 + j4 + j : best_c[j];
      best[j] = fminf(best[j], dj);
    }
  }
}

// v[i] = this lane's partial sum of row i  ->  returns the total of row `lane` (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// the oracle's exact arithmetic for one (row, codeword): sequential fma over the dims.  The row comes from global
// memory (two L1 lines), the codeword from the resident planes (hi + lo is the exact fp32 element).  Inlined and
// streaming: no register array, no call.
template <int K>
__device__ __forceinline__ float exact_dist(const float* __restrict__ zrow, const uint8_t* __restrict__ sB, int k,
                                            float eek) {
  constexpr int B_PLANE = K * 128;
  float zz = 0.f, dot = 0.f;
#pragma unroll 4
  for (int c = 0; c < 16; ++c) {
    const float4 zv = __ldg(reinterpret_cast<const float4*>(zrow) + c);
    const uint8_t* a = sB + ((c >> 3) * 2) * B_PLANE + (uint32_t)k * 128u + (uint32_t)((((c & 7) ^ k) & 7) << 4);
    const float4 hi = *reinterpret_cast<const float4*>(a);
    const float4 lo = *reinterpret_cast<const float4*>(a + B_PLANE);
    const float ev[4] = {hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w};
    const float zc[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      zz = fmaf(zc[j], zc[j], zz);
      dot = fmaf(zc[j], ev[j], dot);
    }
  }
  return (zz - 2.f * dot) + eek;
}

// Warp roles: warps 0-3 and 4-7 two epilogue groups (group g takes the CTA's tiles it = g, g + 2, ... and TMEM
// accumulator buffer g, so two tiles are in flight per SM), warps 8-9 operand staging, warps 10-11 head sum, warp 12
// MMA issue.
// Pipeline per cluster (one CTA per head, every CTA walks the same row tiles t = blockIdx.x, + gridDim.x, ...):
//   stagers   : A(t) -> shared memory (single stage, released by the tile's last MMA)            -> a_full
//   MMA warp  : 24 tcgen05.mma into accumulator buffer t & 1 (2 x K TMEM columns)                -> a_free, acc_full[b]
//   epilogue  : (lanes = dims) load its 32 rows coalesced, |z|^2 by warp reduction;
//               (thread = row = TMEM lane) scan the K scores, pick the index                       -> acc_free[b]
//               (lanes = dims) per row: codeword from the resident planes (one conflict-free 128-byte row per
//               half and plane; hi + lo is the exact fp32 value), quant_raw / quant_st as full 128-byte lines,
//               (q - z)^2 PUSHED into the inbox of the CTA that owns the row (rows [o*R, (o+1)*R) of a tile belong
//               to CTA o, R = 128 / heads)                                                         -> inbox_full
//   stagers   : (after staging tile t+1) sum the inbox over heads in head order, write diff       -> inbox_free
// so the loads of tile t+1 and its MMAs run under the epilogue of tile t, every global access is a full line, and the
// only cross-CTA traffic is 24 KB of fire-and-forget remote stores plus remote mbarrier arrivals -- no cluster-wide
// barrier inside the loop.
template <int K>   // codewords per head: 64, 128 or 256 (= the MMA's N)
__global__ void __launch_bounds__(VU_THREADS, 1)
vq_search_umma_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ embed,
                      float* __restrict__ quant_raw, float* __restrict__ quant_st, float* __restrict__ diff,
                      int64_t* __restrict__ idx, int n_rows) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_PLANE = VU_ROWS * 128;         // 16 KB: 128 rows x 32 dims
  constexpr int B_PLANE = K * 128;               // K codewords x 32 dims
  // [half][plane]: half = dims 0..31 / 32..63, plane = hi / lo
  uint8_t* sA = smem;                                        // 4 x 16 KB
  uint8_t* sB = sA + 4 * A_PLANE;                            // 4 x B_PLANE
  float* inbox = reinterpret_cast<float*>(sB + 4 * B_PLANE); // [head][R rows][64 dims] : 32 KB
  float* ee = inbox + VU_ROWS * VU_DIM;                      // [K] exact |e_k|^2
  float* ee_max = ee + K;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ee_max + 2);
  uint64_t* a_full = bars;            // count 2   (staging warps)
  uint64_t* a_free = bars + 1;        // count 1   (tcgen05.commit)
  uint64_t* acc_full = bars + 2;      // [2] count 1 (tcgen05.commit)
  uint64_t* acc_free = bars + 4;      // [2] count 4 (epilogue warps)
  uint64_t* inbox_full = bars + 6;    // count 1 (the owner's expect_tx) + 32 KB of st.async bytes per tile
  uint64_t* inbox_free = bars + 7;    // [2] (tile parity) count n_heads (one remote arrival per owner CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_heads = gridDim.y, h = blockIdx.y;
  const int lgR = 7 - (31 - __clz(n_heads));     // R = 128 / n_heads rows of a tile owned by each CTA (heads 1/2/4/8)
  const int R = 1 << lgR;
  const int n_tiles = (n_rows + VU_ROWS - 1) / VU_ROWS;
  const float* e_h = embed + (size_t)h * VU_DIM * K;
  constexpr int MMA_WARP = (VU_EPI + VU_STAGERS) / 32;

  if (tid == 0) {
    mbar_init(a_full, 2);
    mbar_init(a_free, 1);
    mbar_init(acc_full, 1); mbar_init(acc_full + 1, 1);
    mbar_init(acc_free, 4); mbar_init(acc_free + 1, 4);
    mbar_init(inbox_full, 1);
    mbar_arrive_expect_tx(inbox_full, VU_ROWS * VU_DIM * 4);     // tile 0: every head pushes its R rows x 64 dims
    mbar_init(inbox_free, (uint32_t)n_heads); mbar_init(inbox_free + 1, (uint32_t)n_heads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();        // barriers initialised: the staging warps may signal a_full during the prologue
  // A operand of one tile: 128 rows x 16 sixteen-byte chunks; consecutive threads read consecutive chunks of a row
  // (256 B per row); hi / lo planes, K-major SWIZZLE_128B
  // `nthr` threads (st = 0 .. nthr-1) share the tile; eight loads are issued before the first is consumed (the
  // compiler did not unroll this loop inside the lambda: 32 dependent round trips made a single-tile launch 10 us
  // longer than the pipelined per-tile period)
  auto stage_rows = [&](int tile, int st, int nthr) {
    const int row0 = tile * VU_ROWS;
    for (int e0 = st; e0 < VU_ROWS * 16; e0 += 8 * nthr) {
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + j * nthr;
        const int r = e >> 4, c_all = e & 15;
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n_rows)
          x[j] = __ldg(reinterpret_cast<const float4*>(z + (int64_t)(row0 + r) * ld_z + h * VU_DIM + c_all * 4));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + j * nthr;
        const int r = e >> 4, c_all = e & 15;
        const int half = c_all >> 3, chunk = c_all & 7;
        const float4 hi = make_float4(tf32_hi(x[j].x), tf32_hi(x[j].y), tf32_hi(x[j].z), tf32_hi(x[j].w));
        uint8_t* d = sA + (half * 2) * A_PLANE + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
                     (uint32_t)((chunk ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(d) = hi;
        *reinterpret_cast<float4*>(d + A_PLANE) =
            make_float4(x[j].x - hi.x, x[j].y - hi.y, x[j].z - hi.z, x[j].w - hi.w);
      }
    }
  };
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * K))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp < VU_EPI / 32) {
    // ============ once per CTA: this head's codebook -> hi / lo operand planes, exact norms, max norm ============
    // lane = (codeword kk of a block of 8, one of 4 sixteen-byte dim chunks): the global reads are full 32-byte
    // sectors along the codewords of one dim (the codebook is dim-major), the 16-byte shared-memory stores of a
    // quarter warp hit the 8 distinct chunk positions c ^ (k & 7) of 8 consecutive rows: conflict-free
    {
      // (K / 8) * 4 items over 8 warps, 8 items = 32 loads in flight per thread and round trip: the prologue is a
      // chain of dependent L2 / DRAM round trips and a single-tile launch is mostly prologue
      const int kk = lane & 7, dc = lane >> 3;
      constexpr int ITEMS = (K / 8) * 4 / (VU_EPI / 32);      // items per warp: 4, 8 or 16
      constexpr int BATCH = ITEMS < 8 ? ITEMS : 8;
#pragma unroll 1
      for (int i0 = 0; i0 < ITEMS; i0 += BATCH) {
        float xv[BATCH][4];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          const int item = warp + (VU_EPI / 32) * (i0 + b);
          const int k = (item >> 2) * 8 + kk, c = (item & 3) * 4 + dc;    // codeword, 16-byte chunk of its 64 dims
#pragma unroll
          for (int j = 0; j < 4; ++j) xv[b][j] = __ldg(e_h + (size_t)(c * 4 + j) * K + k);
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          const int item = warp + (VU_EPI / 32) * (i0 + b);
          const int k = (item >> 2) * 8 + kk, c = (item & 3) * 4 + dc;
          const float4 hi = make_float4(tf32_hi(xv[b][0]), tf32_hi(xv[b][1]), tf32_hi(xv[b][2]), tf32_hi(xv[b][3]));
          uint8_t* d = sB + ((c >> 3) * 2) * B_PLANE + (uint32_t)k * 128u + (uint32_t)((((c & 7) ^ k) & 7) << 4);
          *reinterpret_cast<float4*>(d) = hi;
          *reinterpret_cast<float4*>(d + B_PLANE) =
              make_float4(xv[b][0] - hi.x, xv[b][1] - hi.y, xv[b][2] - hi.z, xv[b][3] - hi.w);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the planes are read by the async proxy (MMA)
    asm volatile("bar.sync 1, %0;" ::"n"(VU_EPI) : "memory");
    // exact codeword norms in the oracle's order (sequential fma over d), from the resident planes: hi + lo is the
    // exact fp32 element
    for (int k = tid; k < K; k += VU_EPI) {
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const uint8_t* a = sB + ((c >> 3) * 2) * B_PLANE + (uint32_t)k * 128u + (uint32_t)((((c & 7) ^ k) & 7) << 4);
        const float4 hi = *reinterpret_cast<const float4*>(a);
        const float4 lo = *reinterpret_cast<const float4*>(a + B_PLANE);
        float v;
        v = hi.x + lo.x; sq = fmaf(v, v, sq);
        v = hi.y + lo.y; sq = fmaf(v, v, sq);
        v = hi.z + lo.z; sq = fmaf(v, v, sq);
        v = hi.w + lo.w; sq = fmaf(v, v, sq);
      }
      ee[k] = sq;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(VU_EPI) : "memory");
    if (warp == 0) {
      float m = 0.f;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, ee[k]);
      m = warp_max(m);
      if (lane == 0) ee_max[0] = m;
    }
  } else if (warp < MMA_WARP) {
    // the first tile's rows arrive while the codebook is being staged: all four staging / head-sum warps take part
    stage_rows(blockIdx.x, tid - VU_EPI, VU_STAGERS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 3, %0;" ::"n"(VU_STAGERS) : "memory");
    if (warp < VU_EPI / 32 + 2 && lane == 0) mbar_arrive(a_full);      // (count 2: the two staging warps)
  }
  tc_fence_before();
  __syncthreads();
  cluster.sync();         // every CTA's barriers are initialised before any remote arrival / store targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float inv_heads = 1.f / (float)n_heads;

  if (warp < VU_EPI / 32) {
    // ================================================ epilogue ================================================
    const int grp = warp >> 2, quad = warp & 3;          // TMEM lanes 32 * quad .. + 31 (a warp reaches quadrant warp % 4)
    const uint32_t taddr_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)grp * (uint32_t)K;
    const uint8_t* sBl = sB + (uint32_t)(lane & 3) * 4u;
    const int lx = lane >> 2;
    const float emax = ee_max[0];
    const int wr0 = quad * 32;                           // first tile row of this warp
    const int64_t row_pitch = (int64_t)n_heads * VU_DIM;
    // the inbox slots of this warp's rows: rows 0-15 and 16-31 (different owners only at 8 heads, R = 16)
    const uint32_t inbox_addr = smem_u32(inbox), full_addr = smem_u32(inbox_full);
    const uint32_t dst_a = map_to_cta(inbox_addr + (uint32_t)(((h << lgR) + (wr0 & (R - 1))) * VU_DIM + lane) * 4u,
                                      (uint32_t)(wr0 >> lgR));
    const uint32_t dst_b = map_to_cta(inbox_addr + (uint32_t)(((h << lgR) + ((wr0 + 16) & (R - 1))) * VU_DIM + lane) * 4u,
                                      (uint32_t)((wr0 + 16) >> lgR));
    const uint32_t full_a = map_to_cta(full_addr, (uint32_t)(wr0 >> lgR));
    const uint32_t full_b = map_to_cta(full_addr, (uint32_t)((wr0 + 16) >> lgR));
    uint32_t it = (uint32_t)grp;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < n_tiles; tile += 2 * gridDim.x, it += 2) {
      const int row_w = tile * VU_ROWS + wr0;            // global row of this warp's first row
      // ---- lanes = dims: the warp's 32 rows as full 128-byte lines; |z|^2 of row i lands in lane i
      float z0[32], z1[32];
      {
        const float* zp = z + (int64_t)row_w * ld_z + h * VU_DIM + lane;
        const int last = n_rows - 1 - row_w;             // rows past the end re-read the last row (never stored)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float* zi = zp + (int64_t)min(i, last) * ld_z;
          z0[i] = __ldg(zi);
          z1[i] = __ldg(zi + 32);
        }
      }
      float zz;
      {
        float part[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) part[i] = fmaf(z1[i], z1[i], z0[i] * z0[i]);
        zz = warp_transpose_reduce32(part, lane);
      }
      const float delta2 = 2.f * 9.5367431640625e-07f * (zz + emax + 2.f * sqrtf(zz * emax));   // 2 * 2^-20 * (|z| + |e|)^2

      // ---- thread = row: one pass over the K scores (accumulator buffer = group)
      mbar_wait(acc_full + grp, (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = taddr_lane;
      float best, second;
      int best_k;
      {
        float b4[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, s4[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        int k4[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
        uint32_t acc0[16], acc1[16];
        tmem_ld16_issue(taddr, acc0);
#pragma unroll 1
        for (int c0 = 0; c0 < K; c0 += 32) {
          tmem_ld_wait16(acc0);
          tmem_ld16_issue(taddr + (uint32_t)(c0 + 16), acc1);
          scan16(acc0, ee, c0, zz, b4, s4, k4);
          tmem_ld_wait16(acc1);
          if (c0 + 32 < K) tmem_ld16_issue(taddr + (uint32_t)(c0 + 32), acc0);
          scan16(acc1, ee, c0 + 16, zz, b4, s4, k4);
        }
        // merge: the overall minimum, its index, and the smallest value that is not the winner's.  (An exact tie
        // between accumulators leaves second == best: the row is ambiguous and the exact pass below applies the
        // lowest-index rule.)
        best = b4[0]; second = s4[0]; best_k = k4[0];
#pragma unroll
        for (int a = 1; a < 4; ++a) {
          const bool wins = b4[a] < best;
          second = fminf(fminf(second, s4[a]), wins ? best : b4[a]);
          best_k = wins ? k4[a] : best_k;
          best = wins ? b4[a] : best;
        }
      }
      // A row whose runner-up lies more than 2*delta above its best has a single candidate -- the exhaustive fp32
      // search's argmin (tests/test_vq_two_phase_margin.py); any other row is re-scored exactly.
      const bool amb = !(second > best + delta2) || (unsigned)best_k >= (unsigned)K;
      if (__any_sync(0xffffffffu, amb)) {
        // some row of this warp is ambiguous (or NaN / Inf): the warp re-reads its rows' K dot products
        // (tcgen05.ld is warp-collective) and the ambiguous lanes re-score every codeword within 2*delta of their
        // minimum with the oracle's exact arithmetic, in ascending index order (lowest index wins ties)
        const float thr = best + delta2;
        const float* zrow = z + (int64_t)min(row_w + lane, n_rows - 1) * ld_z + h * VU_DIM;
        float bd = INFINITY;
        int bk = 0x7fffffff;
        // A SMALL LOOPED body (32 columns per iteration: two warp-collective TMEM loads, a 32-bit candidate mask,
        // the exact re-score of its set bits in ascending order): this block runs about once per 8 000 row-heads, so
        // its code is never in the instruction cache -- fully unrolled (1 300 instructions of straight-line code) it
        // cost 22 - 25 k cycles of instruction fetch for ~5 k cycles of work, which is what separated the 18 us and
        // the 31 us single-tile launches (profiles/r02_vq_small_probe.txt)
#pragma unroll 1
        for (int w = 0; w < K / 32; ++w) {
          uint32_t m = 0u;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float acc[16];
            tmem_ld16(taddr + (uint32_t)(w * 32 + hf * 16), acc);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((zz - 2.f * acc[j]) + ee[w * 32 + hf * 16 + j] <= thr) m |= 1u << (hf * 16 + j);
          }
          if (!amb) m = 0u;
          while (m) {
            const int k = w * 32 + __ffs(m) - 1;
            m &= m - 1;
            const float dk = exact_dist<K>(zrow, sB, k, ee[k]);
            if (dk < bd) { bd = dk; bk = k; }
          }
        }
        if (amb) best_k = ((unsigned)bk < (unsigned)K) ? bk : 0;     // no candidate at all (NaN row): index 0
      }
      tc_fence_before();                                   // this thread's TMEM reads of the tile are complete
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free + grp);
      if (row_w + lane < n_rows) idx[(int64_t)(row_w + lane) * n_heads + h] = (int64_t)best_k;
      // ---- lanes = dims, four rows in flight: codeword k (dims lane, 32 + lane: one 128-byte row per half and
      //      plane, 16-byte chunks XORed with k & 7); quant_raw / quant_st as full 128-byte lines.  This part needs
      //      nothing from the other CTAs, so it runs BEFORE the wait for the inbox.
      float* qr = quant_raw + (int64_t)row_w * row_pitch + h * VU_DIM + lane;
      float* qs = quant_st + (int64_t)row_w * row_pitch + h * VU_DIM + lane;
      const int n_ok = n_rows - row_w;                     // rows i < n_ok exist
#pragma unroll
      for (int i0 = 0; i0 < 32; i0 += 4) {
        float q0[4], q1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = __shfl_sync(0xffffffffu, best_k, i0 + j);
          const uint8_t* a = sBl + (uint32_t)k * 128u + (uint32_t)(((lx ^ k) & 7) << 4);
          q0[j] = *reinterpret_cast<const float*>(a) + *reinterpret_cast<const float*>(a + B_PLANE);
          q1[j] = *reinterpret_cast<const float*>(a + 2 * B_PLANE) + *reinterpret_cast<const float*>(a + 3 * B_PLANE);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          if (i < n_ok) {                                    // (warp-uniform)
            qr[0] = q0[j];
            qr[32] = q1[j];
            qs[0] = z0[i] + (q0[j] - z0[i]);
            qs[32] = z1[i] + (q1[j] - z1[i]);
          }
          qr += row_pitch;
          qs += row_pitch;
        }
      }
      // the owners have summed the previous tile (the other group's): the inboxes may be overwritten.  The owners
      // signal tile t on inbox_free[t & 1], so each group watches ONE barrier and sees every one of its phases (with
      // a single barrier a group could be 0, 1 or 2 phases behind, which a parity wait cannot tell apart).
      if (it > 0) mbar_wait_cluster(inbox_free + ((it - 1) & 1u), ((it - 1) >> 1) & 1u);
      // ---- push (q - z)^2 of the 32 rows to their owners (gathers the codewords again: 4 LDS per row)
#pragma unroll
      for (int i0 = 0; i0 < 32; i0 += 4) {
        float q0[4], q1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = __shfl_sync(0xffffffffu, best_k, i0 + j);
          const uint8_t* a = sBl + (uint32_t)k * 128u + (uint32_t)(((lx ^ k) & 7) << 4);
          q0[j] = *reinterpret_cast<const float*>(a) + *reinterpret_cast<const float*>(a + B_PLANE);
          q1[j] = *reinterpret_cast<const float*>(a + 2 * B_PLANE) + *reinterpret_cast<const float*>(a + 3 * B_PLANE);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          const float dq0 = q0[j] - z0[i], dq1 = q1[j] - z1[i];
          const uint32_t dst = (i < 16 ? dst_a + (uint32_t)i * 256u : dst_b + (uint32_t)(i - 16) * 256u);
          const uint32_t bar = (i < 16 ? full_a : full_b);
          st_async_f32(dst, __fmul_rn(dq0, dq0), bar);       // no fma contraction with the head sum
          st_async_f32(dst + 128u, __fmul_rn(dq1, dq1), bar);
        }
      }
    }
  } else if (warp < VU_EPI / 32 + 2) {
    // ============================= operand staging (A): runs ahead, gated by a_free =============================
    const int st = tid - VU_EPI;                         // 0..63
    uint32_t it = 1;                                     // (tile 0 was staged during the prologue)
    for (int tile = blockIdx.x + gridDim.x; tile < n_tiles; tile += gridDim.x, ++it) {
      mbar_wait(a_free, (it - 1) & 1u);                    // the previous tile's MMAs have read the stage
      stage_rows(tile, st, 64);
      publish_and_arrive_warp(a_full);
    }
  } else if (warp < MMA_WARP) {
    // ================================ head sum of the commitment term ================================
    // this CTA owns rows [h*R, (h+1)*R) of every tile: sum the heads in head order, scale, write diff
    const int st = tid - VU_EPI - 64;                    // 0..63
    const uint32_t inbox_sa = smem_u32(inbox);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * VU_ROWS;
      mbar_wait_cluster(inbox_full, it & 1u);
      if (st == 0) mbar_arrive_expect_tx(inbox_full, VU_ROWS * VU_DIM * 4);    // the next tile's bytes
      // sum of one float4 (row lr, dims 4*d4 ..) over the heads in head order; the loads of all heads are issued
      // before the first add (the head count is a runtime value: predicated, fully unrolled)
      auto sum_item = [&](int e) {
        const uint32_t a0 = inbox_sa + (uint32_t)(((e >> 4) * VU_DIM + (e & 15) * 4) * 4);
        float4 v[8];
#pragma unroll
        for (int hh = 0; hh < 8; ++hh)
          if (hh < n_heads)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[hh].x), "=f"(v[hh].y), "=f"(v[hh].z), "=f"(v[hh].w)
                         : "r"(a0 + (uint32_t)((hh << lgR) * VU_DIM * 4)));
        float4 a = v[0];
#pragma unroll
        for (int hh = 1; hh < 8; ++hh)
          if (hh < n_heads) {
            a.x = __fadd_rn(a.x, v[hh].x); a.y = __fadd_rn(a.y, v[hh].y);
            a.z = __fadd_rn(a.z, v[hh].z); a.w = __fadd_rn(a.w, v[hh].w);
          }
        return make_float4(a.x * inv_heads, a.y * inv_heads, a.z * inv_heads, a.w * inv_heads);
      };
      auto store_item = [&](int e, const float4& v) {
        const int row = row0 + (h << lgR) + (e >> 4);
        if (row < n_rows) *reinterpret_cast<float4*>(diff + (int64_t)row * VU_DIM + (e & 15) * 4) = v;
      };
      if (R <= 32) {
        // (>= 4 heads) the sums wait in registers: the inbox is handed back before the global stores are issued
        float4 acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (st + 64 * u < R * 16) acc[u] = sum_item(st + 64 * u);
        asm volatile("bar.sync 2, 64;" ::: "memory");      // both warps hold their sums: the inbox is free
        if (st < n_heads) mbar_arrive_remote_release(map_to_cta(smem_u32(inbox_free + (it & 1u)), (uint32_t)st));
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (st + 64 * u < R * 16) store_item(st + 64 * u, acc[u]);
      } else {
        for (int e = st; e < R * 16; e += 64) store_item(e, sum_item(e));
        asm volatile("bar.sync 2, 64;" ::: "memory");
        if (st < n_heads) mbar_arrive_remote_release(map_to_cta(smem_u32(inbox_free + (it & 1u)), (uint32_t)st));
      }
    }
  } else {
    // ================================ MMA issuer (convergent, predicated issue) ================================
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) |                    // D f32, A/B tf32, both K-major
                               ((uint32_t)(K >> 3) << 17) | ((uint32_t)(VU_ROWS >> 4) << 24);
    const uint32_t elected = elect_one();
    const uint32_t a0 = desc_lo_k(smem_u32(sA)), b0 = desc_lo_k(smem_u32(sB));
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1u;
      if (it >= 2) mbar_wait(acc_free + b, ((it >> 1) - 1u) & 1u);   // epilogue group b has drained its accumulator
      mbar_wait(a_full, it & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + b * (uint32_t)K;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t ad = a0 + (uint32_t)(half * 2) * (A_PLANE >> 4);
        const uint32_t bd = b0 + (uint32_t)(half * 2) * (B_PLANE >> 4);
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          // 8 dims per MMA: both operands advance 32 B along their 128-byte rows (2 descriptor units)
          const uint32_t a_hi = ad + 2 * kg, b_hi = bd + 2 * kg;
          const uint32_t a_lo = a_hi + (A_PLANE >> 4), b_lo = b_hi + (B_PLANE >> 4);
          umma_tf32_pred<DESC_HI_K>(tmem_d, a_lo, b_hi, IDESC, (half > 0 || kg > 0) ? 1u : 0u, elected);
          umma_tf32_pred<DESC_HI_K>(tmem_d, a_hi, b_lo, IDESC, 1u, elected);
          umma_tf32_pred<DESC_HI_K>(tmem_d, a_hi, b_hi, IDESC, 1u, elected);
        }
      }
      umma_commit_pred(a_free, elected);
      umma_commit_pred(acc_full + b, elected);
      __syncwarp();
    }
  }
  // nobody exits (or frees its TMEM) while a neighbour may still push into its inbox / arrive on its barriers
  tc_fence_before();
  __syncthreads();
  cluster.sync();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * K))
                 : "memory");
  }
}

template <int K>
int launch_vq_umma(const float* z, int64_t ld_z, const float* embed, float* quant_raw, float* quant_st, float* diff,
                   int64_t* idx, int n_rows, int n_heads, cudaStream_t st) {
  const size_t smem = 1024 + 4 * (size_t)VU_ROWS * 128 + 4 * (size_t)K * 128 + (size_t)VU_ROWS * VU_DIM * 4 +
                      (size_t)K * 4 + 8 + 8 * 8 + 16;
  static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // per n_heads: co-resident clusters of this kernel
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(VU_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)n_heads;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters[n_heads] == 0) {
    if (cudaFuncSetAttribute(vq_search_umma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return MSMC_ERR_LAUNCH;
    // persistent: the grid must not exceed what is co-resident (clusters are placed inside one GPC, so this can be
    // fewer than num_sms / n_heads); a cluster left for a second wave would double the kernel's duration
    cfg.gridDim = dim3((unsigned)(num_sms() / n_heads), (unsigned)n_heads, 1);
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, vq_search_umma_kernel<K>, &cfg) != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      nc = std::max(1, num_sms() / n_heads - 2);
    }
    max_clusters[n_heads] = std::min(nc, num_sms() / n_heads);
  }
  int n_clusters = std::max(1, std::min(ceil_div(n_rows, VU_ROWS), max_clusters[n_heads]));
  if (const char* e = getenv("MSMC_VQ_MAX_CLUSTERS")) n_clusters = std::max(1, std::min(n_clusters, atoi(e)));
  cfg.gridDim = dim3((unsigned)n_clusters, (unsigned)n_heads, 1);
  if (cudaLaunchKernelEx(&cfg, vq_search_umma_kernel<K>, z, ld_z, embed, quant_raw, quant_st, diff, idx, n_rows) !=
      cudaSuccess)
    return MSMC_ERR_LAUNCH;
  return MSMC_OK;
}

}  // namespace
}  // namespace msmc

using namespace msmc;

extern "C" int msmc_vq_search_umma(const float* z, int64_t ld_z, const float* embed, float* quant_raw,
                                   float* quant_st, float* diff, int64_t* idx, int32_t n_rows, int32_t n_heads,
                                   int32_t dim, int32_t n_embed, void* stream) {
  MSMC_REQUIRE(z && embed && quant_raw && quant_st && diff && idx);
  MSMC_REQUIRE(n_rows > 0 && n_heads > 0 && n_heads <= 8);
  if (n_heads != 1 && n_heads != 2 && n_heads != 4 && n_heads != 8) return MSMC_ERR_UNSUPPORTED;
  if (dim != VU_DIM || (n_embed != 64 && n_embed != 128 && n_embed != 256)) return MSMC_ERR_UNSUPPORTED;
  MSMC_REQUIRE((ld_z & 3) == 0 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(embed) |
                                    reinterpret_cast<uintptr_t>(quant_raw) | reinterpret_cast<uintptr_t>(quant_st)) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (n_embed == 64) rc = launch_vq_umma<64>(z, ld_z, embed, quant_raw, quant_st, diff, idx, n_rows, n_heads, st);
  else if (n_embed == 128) rc = launch_vq_umma<128>(z, ld_z, embed, quant_raw, quant_st, diff, idx, n_rows, n_heads, st);
  else rc = launch_vq_umma<256>(z, ld_z, embed, quant_raw, quant_st, diff, idx, n_rows, n_heads, st);
  if (rc != MSMC_OK) return rc;
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
