// Two-phase multi-head VQ search on the tensor cores (msmc_vq_search_umma; validated bit-exact against
// oracle/vq_oracle.c by tests/test_vq_umma_gpu.py).
//
// The CUDA-core search kernels (vq.cu) are fp32-FMA bound: at K = 256 a row costs 65 536 FMAs per 3 360 bytes, so the
// exhaustive search tops out near 28 % of the HBM roofline.  Here
//   phase 1  scores a 128-row tile of one head against all K codewords with tcgen05.mma (3xTF32: fp32-accurate to
//            ~2^-21 relative, accumulators in TMEM): A = z rows, K-major SWIZZLE_128B (like conv_umma_kernel);
//            B = the head's codebook straight from its dim-major `embed` buffer, MN-major SWIZZLE_128B_BASE32B
//            (like the weight-gradient kernel's operands: a 128-byte shared-memory row = 32 consecutive codewords of
//            one dim), two K halves of 32 dims;
//   phase 2  reads the 128 x K dot products from TMEM, forms dist = (|z|^2 - 2 z.e_k) + |e_k|^2 with the exact
//            (sequential-fma) norms, and keeps every codeword within 2*delta of the row minimum,
//            delta = 2^-20 (|z|^2 + max|e|^2 + 2 |z| max|e|) (tests/test_vq_two_phase_margin.py: the exhaustive search's
//            argmin is always in that set and the set is a singleton for > 99.9 % of the rows).  Rows with more than
//            one candidate re-score the candidates with the oracle's exact arithmetic and tie rule (lowest index).
// The result is identical to the exhaustive fp32 search by construction.
// PERSISTENT: the heads of a row tile form a thread-block cluster (CTA = head) and a cluster walks the row tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ...  Each CTA stages ITS head's codebook (hi / lo planes, exact norms) in
// shared memory ONCE and keeps it for every tile (the one-tile-per-CTA form re-read 64 KB of codebook per 128 rows:
// more L2 traffic than the rows themselves).  The commitment term's head sum goes through distributed shared memory
// in head order, as in vq_search_cluster_kernel.
#include "umma.cuh"
#include <algorithm>
#include <cooperative_groups.h>
#include <cstdlib>

namespace msmc {
namespace {

constexpr int VU_ROWS = 128;                 // rows per CTA (TMEM lanes)
constexpr int VU_DIM = 64;                   // dims per head (two K halves of 32)
constexpr int VU_PRODUCERS = 256;            // 8 producer / epilogue warps
constexpr int VU_THREADS = VU_PRODUCERS + 32;  // + the MMA warp
constexpr int VU_MAX_CAND = 4;               // candidates kept per row half before the exhaustive fallback
constexpr int VU_DV_LD = VU_DIM + 1;         // padded row pitch of the (q - z)^2 tile: a warp's 32 rows hit 32 banks

// A and B planes use different shared-memory layouts, hence different descriptor high words; convergent predicated
// issue (see umma_tf32_pred in umma.cuh)
__device__ __forceinline__ void umma_tf32_kmaj_a_mnmaj_b_pred(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32,
                                                              uint32_t idesc, uint32_t accumulate, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "n"(DESC_HI_K), "n"(DESC_HI_MN),
        "r"(elected)
      : "memory");
}

// the oracle's exact distance of row z (64 dims, registers) to codeword k of head codebook e (dim-major, ld = K)
__device__ __forceinline__ float exact_dist(const float (&zr)[VU_DIM], float zz, const float* __restrict__ e, int K, int k,
                                            float eek) {
  float dot = 0.f;
#pragma unroll
  for (int d = 0; d < VU_DIM; ++d) dot = fmaf(zr[d], __ldg(e + (size_t)d * K + k), dot);   // (zr stays in registers)
  return (zz - 2.f * dot) + eek;
}

// exact fp32 codeword element (dim d, codeword k) from the resident operand planes: hi + lo is exact by construction
template <int K>
__device__ __forceinline__ float cb_elem(const uint8_t* __restrict__ sB, int d, int k) {
  constexpr int B_PLANE = (K / 32) * 4096;
  const int half = d >> 5, p = d & 31, c = k >> 2;
  const uint8_t* a = sB + (half * 2) * B_PLANE + (uint32_t)(c >> 3) * 4096u + mn_off(p, c & 7) + (uint32_t)(k & 3) * 4u;
  return *reinterpret_cast<const float*>(a) + *reinterpret_cast<const float*>(a + B_PLANE);
}

// dims D0 .. D0+31 of one row: gather the chosen codeword (from shared memory), write quant_raw / quant_st, leave
// (q - z)^2 in `dv`
template <int K, int D0>
__device__ __forceinline__ void emit_row_half(const float (&zr)[VU_DIM], const uint8_t* __restrict__ sB, int k,
                                              bool row_ok, int row, int r, int n_heads, int h,
                                              float* __restrict__ quant_raw, float* __restrict__ quant_st,
                                              float* __restrict__ dv) {
  float q[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) q[j] = cb_elem<K>(sB, D0 + j, k);
  if (row_ok) {
    float* qr = quant_raw + (int64_t)row * (n_heads * VU_DIM) + h * VU_DIM + D0;
    float* qs = quant_st + (int64_t)row * (n_heads * VU_DIM) + h * VU_DIM + D0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      *reinterpret_cast<float4*>(qr + j) = make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]);
      float4 st;
      st.x = zr[D0 + j] + (q[j] - zr[D0 + j]);
      st.y = zr[D0 + j + 1] + (q[j + 1] - zr[D0 + j + 1]);
      st.z = zr[D0 + j + 2] + (q[j + 2] - zr[D0 + j + 2]);
      st.w = zr[D0 + j + 3] + (q[j + 3] - zr[D0 + j + 3]);
      *reinterpret_cast<float4*>(qs + j) = st;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float dq = q[j] - zr[D0 + j];
    dv[r * VU_DV_LD + D0 + j] = __fmul_rn(dq, dq);    // no fma contraction with the head sum
  }
}

template <int K>   // codewords per head: 64, 128 or 256 (= the MMA's N)
__global__ void __launch_bounds__(VU_THREADS, 1)
vq_search_umma_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ embed,
                      float* __restrict__ quant_raw, float* __restrict__ quant_st, float* __restrict__ diff,
                      int64_t* __restrict__ idx, int n_rows) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NB = K / 32;                     // 32-codeword blocks of the B operand
  constexpr int A_PLANE = VU_ROWS * 128;         // 16 KB: 128 rows x 32 dims
  constexpr int B_PLANE = NB * 4096;             // NB blocks x 32 dims x 128 B
  // [half][plane]: half = dims 0..31 / 32..63, plane = hi / lo
  uint8_t* sA = smem;                                        // 4 x 16 KB
  uint8_t* sB = sA + 4 * A_PLANE;                            // 4 x B_PLANE
  float* ee = reinterpret_cast<float*>(sB + 4 * B_PLANE);    // [K] exact |e_k|^2
  float* row_best = ee + K;                                  // [2][128] per column half: best approximate distance
  int* cand_cnt = reinterpret_cast<int*>(row_best + 2 * VU_ROWS);   // [2][128]
  int* cand_idx = cand_cnt + 2 * VU_ROWS;                    // [2][128][VU_MAX_CAND]
  int* row_idx = cand_idx + 2 * VU_ROWS * VU_MAX_CAND;       // [128] final index
  float* ee_max = reinterpret_cast<float*>(row_idx + VU_ROWS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ee_max + 2);
  uint64_t* accum_bar = full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* dv = reinterpret_cast<float*>(sA);      // [128][65] per-row (q - z)^2 of this head; aliases A after the MMAs

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_heads = gridDim.y, h = blockIdx.y;
  const int n_tiles = (n_rows + VU_ROWS - 1) / VU_ROWS;
  const float* e_h = embed + (size_t)h * VU_DIM * K;
  constexpr int MMA_WARP = VU_PRODUCERS / 32;

  if (tid == 0) {
    mbar_init(full_bar, VU_PRODUCERS / 32);
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)K)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < MMA_WARP) {
    // ============ once per CTA: this head's codebook -> hi / lo operand planes, exact norms, max norm ============
    // B: 64 dims x K/4 chunks of 4 consecutive codewords, read straight from the dim-major codebook
#pragma unroll 4
    for (int i = 0; i < (VU_DIM * (K / 4)) / VU_PRODUCERS; ++i) {
      const int e = tid + VU_PRODUCERS * i;
      const int d_ = e / (K / 4), c = e - d_ * (K / 4);
      const int half = d_ >> 5, p = d_ & 31, nblk = c >> 3, c16 = c & 7;
      const float4 x = __ldg(reinterpret_cast<const float4*>(e_h + (size_t)d_ * K + c * 4));
      const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
      uint8_t* d = sB + (half * 2) * B_PLANE + (uint32_t)nblk * 4096u + mn_off(p, c16);
      *reinterpret_cast<float4*>(d) = hi;
      *reinterpret_cast<float4*>(d + B_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    }
    // exact codeword norms, the oracle's order (sequential fma over d)
    for (int k = tid; k < K; k += VU_PRODUCERS) {
      float s = 0.f;
#pragma unroll 8
      for (int d_ = 0; d_ < VU_DIM; ++d_) { const float v = __ldg(e_h + (size_t)d_ * K + k); s = fmaf(v, v, s); }
      ee[k] = s;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(VU_PRODUCERS) : "memory");
    if (warp == 0) {
      float m = 0.f;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, ee[k]);
      m = warp_max(m);
      if (lane == 0) ee_max[0] = m;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(VU_PRODUCERS) : "memory");      // ee_max visible
  }
  const float inv_heads = 1.f / (float)n_heads;
  const uint32_t elected = elect_one();

  uint32_t par = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, par ^= 1u) {
  const int row0 = tile * VU_ROWS;
  if (warp < MMA_WARP) {
    // ================================ operand staging: this tile's rows ================================
    // A: 128 rows x 16 sixteen-byte chunks; consecutive threads read consecutive chunks of a row (256 B per row)
#pragma unroll
    for (int i = 0; i < (VU_ROWS * 16) / VU_PRODUCERS; ++i) {
      const int e = tid + VU_PRODUCERS * i;
      const int r = e >> 4, c_all = e & 15;
      const int half = c_all >> 3, chunk = c_all & 7;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < n_rows)
        x = __ldg(reinterpret_cast<const float4*>(z + (int64_t)(row0 + r) * ld_z + h * VU_DIM + c_all * 4));
      const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
      uint8_t* d = sA + (half * 2) * A_PLANE + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
                   (uint32_t)((chunk ^ (r & 7)) << 4);
      *reinterpret_cast<float4*>(d) = hi;
      *reinterpret_cast<float4*>(d + A_PLANE) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    }
    publish_and_arrive_warp(full_bar);      // (on the first tile this also publishes the codebook planes)

    // ================================ phase 2: one thread per (row, column half) ================================
    const int lane_grp = warp & 3, chalf = warp >> 2;
    const int r = lane_grp * 32 + lane;                 // TMEM lane = row inside the tile
    const int row = row0 + r;
    const bool row_ok = row < n_rows;
    const float* zrow = z + (int64_t)(row_ok ? row : (n_rows - 1)) * ld_z + h * VU_DIM;
    float zr[VU_DIM];
#pragma unroll
    for (int d_ = 0; d_ < VU_DIM; d_ += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(zrow + d_));
      zr[d_] = t.x; zr[d_ + 1] = t.y; zr[d_ + 2] = t.z; zr[d_ + 3] = t.w;
    }
    float zz = 0.f;
#pragma unroll
    for (int d_ = 0; d_ < VU_DIM; ++d_) zz = fmaf(zr[d_], zr[d_], zz);
    const float emax = ee_max[0];
    const float delta2 = 2.f * 9.5367431640625e-07f * (zz + emax + 2.f * sqrtf(zz * emax));   // 2 * 2^-20 * (|z| + |e|)^2

    mbar_wait(accum_bar, par);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    constexpr int CH = K / 2;                           // columns per thread
    const int cbeg = chalf * CH;
    // ONE pass over this thread's K/2 columns: smallest approximate distance (lowest index on ties), its index, and
    // the second smallest value.  A row whose runner-up (over both halves) lies more than 2*delta above its best has a
    // single candidate -- the exhaustive fp32 search's argmin (tests/test_vq_two_phase_margin.py); any other row is
    // re-scored exactly over all K codewords (< 0.1 % of the rows).
    float best = INFINITY, second = INFINITY;
    int best_c = 0x7fffffff;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CH; c0 += 16) {
      float acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float dj = (zz - 2.f * acc[j]) + ee[c0 + j];
        if (dj < best) { second = best; best = dj; best_c = c0 + j; }
        else second = fminf(second, dj);
      }
    }
    row_best[chalf * VU_ROWS + r] = best;
    cand_cnt[chalf * VU_ROWS + r] = best_c;
    reinterpret_cast<float*>(cand_idx)[chalf * VU_ROWS + r] = second;
    asm volatile("bar.sync 1, %0;" ::"n"(VU_PRODUCERS) : "memory");

    if (chalf == 0) {                                    // (warp-uniform) the row's owner merges the two halves
      const float b0 = row_best[r], b1 = row_best[VU_ROWS + r];
      const float s0 = reinterpret_cast<const float*>(cand_idx)[r], s1 = reinterpret_cast<const float*>(cand_idx)[VU_ROWS + r];
      const int k0 = cand_cnt[r], k1 = cand_cnt[VU_ROWS + r];
      const bool lo_wins = b0 <= b1;                     // ties: the lower half holds the lower indices
      const float bmin = lo_wins ? b0 : b1;
      const float runner = fminf(lo_wins ? b1 : b0, fminf(s0, s1));
      int best_k = lo_wins ? k0 : k1;
      const bool amb = !(runner > bmin + delta2) || (unsigned)best_k >= (unsigned)K;
      if (__any_sync(0xffffffffu, amb)) {
        // some row of this warp is ambiguous (or NaN / Inf): the warp re-reads its rows' K dot products
        // (tcgen05.ld is warp-collective) and the ambiguous lanes re-score every codeword within 2*delta of their
        // minimum with the oracle's exact arithmetic, in ascending index order (lowest index wins ties)
        const float thr = bmin + delta2;
        float bd = INFINITY;
        int bk = 0x7fffffff;
#pragma unroll 1
        for (int c0 = 0; c0 < K; c0 += 16) {
          float acc[16];
          tmem_ld16(taddr + (uint32_t)c0, acc);
          if (amb) {
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
              if ((zz - 2.f * acc[j]) + ee[c0 + j] <= thr) {
                const float dk = exact_dist(zr, zz, e_h, K, c0 + j, ee[c0 + j]);
                if (dk < bd) { bd = dk; bk = c0 + j; }
              }
            }
          }
        }
        if (amb) best_k = ((unsigned)bk < (unsigned)K) ? bk : 0;     // no candidate at all (NaN row): index 0
      }
      row_idx[r] = best_k;
      if (row_ok) idx[(int64_t)row * n_heads + h] = (int64_t)best_k;
    }
    tc_fence_before();                                   // this thread's TMEM reads of the tile are complete
    asm volatile("bar.sync 1, %0;" ::"n"(VU_PRODUCERS) : "memory");
    // gather, straight-through output and per-row squares: the two threads of a row take 32 dims each.
    // (dv aliases the A operand: every MMA that read it has completed -- accum_bar -- and all threads are past it)
    // (sB is never overwritten; dv aliases sA, so the codeword is gathered into registers BEFORE dv is written --
    //  emit_row_half reads sB only)
    if (chalf == 0)
      emit_row_half<K, 0>(zr, sB, row_idx[r], row_ok, row, r, n_heads, h, quant_raw, quant_st, dv);
    else
      emit_row_half<K, 32>(zr, sB, row_idx[r], row_ok, row, r, n_heads, h, quant_raw, quant_st, dv);
  } else {
    // ================================ MMA issuer (convergent, predicated issue) ================================
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) |      // D f32, A/B tf32, B MN-major
                               ((uint32_t)(K >> 3) << 17) | ((uint32_t)(VU_ROWS >> 4) << 24);
    mbar_wait(full_bar, par);
    tc_fence_after();
    const uint32_t a0 = desc_lo_k(smem_u32(sA)), b0 = desc_lo_mn(smem_u32(sB));
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint32_t ad = a0 + (uint32_t)(half * 2) * (A_PLANE >> 4);
      const uint32_t bd = b0 + (uint32_t)(half * 2) * (B_PLANE >> 4);
#pragma unroll
      for (int kg = 0; kg < 4; ++kg) {
        // 8 dims per MMA: A advances 32 B along its 128-byte rows (2 units), B two 4-row atoms (1 KB = 64 units)
        const uint32_t a_hi = ad + 2 * kg, b_hi = bd + 64 * kg;
        const uint32_t a_lo = a_hi + (A_PLANE >> 4), b_lo = b_hi + (B_PLANE >> 4);
        umma_tf32_kmaj_a_mnmaj_b_pred(tmem_base, a_lo, b_hi, IDESC, (half > 0 || kg > 0) ? 1u : 0u, elected);
        umma_tf32_kmaj_a_mnmaj_b_pred(tmem_base, a_hi, b_lo, IDESC, 1u, elected);
        umma_tf32_kmaj_a_mnmaj_b_pred(tmem_base, a_hi, b_hi, IDESC, 1u, elected);
      }
    }
    umma_commit_pred(accum_bar, elected);
    __syncwarp();
  }
  // head sum of the commitment term in head order through distributed shared memory (cluster = the heads of this tile)
  __syncthreads();
  cluster.sync();
  const int rows_here = min(VU_ROWS, n_rows - row0);
  // this CTA combines the rows ri = h, h + n_heads, ...
  const int my_rows = (rows_here - h + n_heads - 1) / n_heads;
  for (int e = tid; e < my_rows * VU_DIM; e += VU_THREADS) {
    const int ri = (e >> 6) * n_heads + h, d_ = e & (VU_DIM - 1);
    float acc = *cluster.map_shared_rank(dv + ri * VU_DV_LD + d_, 0);
    for (int hh = 1; hh < n_heads; ++hh) acc = __fadd_rn(acc, *cluster.map_shared_rank(dv + ri * VU_DV_LD + d_, hh));
    diff[(int64_t)(row0 + ri) * VU_DIM + d_] = acc * inv_heads;
  }
  cluster.sync();     // nobody overwrites its tile (next A staging) or exits while a neighbour may still read it
  }   // tile loop
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)K) : "memory");
  }
}

template <int K>
int launch_vq_umma(const float* z, int64_t ld_z, const float* embed, float* quant_raw, float* quant_st, float* diff,
                   int64_t* idx, int n_rows, int n_heads, cudaStream_t st) {
  constexpr int NB = K / 32;
  const size_t smem = 1024 + 4 * (size_t)VU_ROWS * 128 + 4 * (size_t)NB * 4096 + (size_t)K * 4 +
                      2 * VU_ROWS * 4 + 2 * VU_ROWS * 4 + 2 * VU_ROWS * VU_MAX_CAND * 4 + VU_ROWS * 4 + 8 + 16 + 16;
  cudaLaunchConfig_t cfg = {};
  // persistent: one cluster (n_heads CTAs, one per SM) per slot, each walking its share of the row tiles
  const int n_clusters = std::max(1, std::min(ceil_div(n_rows, VU_ROWS), num_sms() / n_heads));
  cfg.gridDim = dim3((unsigned)n_clusters, (unsigned)n_heads, 1);
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
  if (cudaFuncSetAttribute(vq_search_umma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return MSMC_ERR_LAUNCH;
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
