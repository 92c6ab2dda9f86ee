// Multi-head VQ: nearest-codeword search, EMA codebook update, straight-through backward, triplet loss.
// Replaces Quantize.forward / MultiHeadQuantize.forward / Quantize.compute_triple_loss
// (reference vqgantts/modules.py:24-67, 86-116, 137-169).
//
// Search kernel: the head's codebook (dim x K, dim-major, exactly the reference's `embed` buffer) is staged in
// shared memory once per CTA together with ||e_k||^2; one warp scores 4 rows at a time, lane l holding 4*KQ
// consecutive-by-4 codewords (one LDS.128 feeds 16 FMAs) with the rows broadcast by warp shuffle, and the argmin is
// a warp-shuffle reduction (lowest index wins ties).  Distances use the reference's expanded form (||z||^2 - 2 z.e_k) + ||e_k||^2 in true fp32 with a
// fixed, sequential fma order that the C oracle (oracle/vq_oracle.c) reproduces bit for bit.
// HBM-bound integer/float gather work: no tensor cores.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cooperative_groups.h>

namespace msmc {
namespace {

constexpr int VQ_WARPS = 8;

// VEC codewords per lane per 16-byte shared-memory load, KQ such loads per dim: the CTA's padded codebook width is
// Kp = 32 * VEC * KQ >= K (padding codewords get distance +inf).  Each warp scores R = 4 rows at once so one
// LDS.128 of the codebook feeds 16 FMAs.
constexpr int VQ_R = 4;

template <int DIM, int VEC, int KQ>
__global__ void __launch_bounds__(VQ_WARPS * 32)
vq_search_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ embed,
                 float* __restrict__ quant_raw, float* __restrict__ quant_st, float* __restrict__ diff,
                 int64_t* __restrict__ idx, int n_rows, int n_heads, int K, int rows_per_cta) {
  extern __shared__ __align__(16) float smem[];
  constexpr int KP = 32 * VEC * KQ;
  float* cb = smem;                      // [DIM][KP]
  float* ee = smem + (size_t)DIM * KP;   // [KP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_beg = blockIdx.x * rows_per_cta;
  const int row_end = min(n_rows, row_beg + rows_per_cta);
  const float inv_heads = 1.f / (float)n_heads;
  constexpr int DPL = DIM / 32;  // dims per lane

  for (int h = 0; h < n_heads; ++h) {
    __syncthreads();  // previous head's codebook fully consumed
    const float* e_h = embed + (size_t)h * DIM * K;
    for (int i = threadIdx.x; i < DIM * KP; i += blockDim.x) {
      const int d = i / KP, k = i - d * KP;
      cb[i] = (k < K) ? e_h[(size_t)d * K + k] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < KP; k += blockDim.x) {
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < DIM; ++d) { float v = cb[d * KP + k]; s = fmaf(v, v, s); }
      ee[k] = (k < K) ? s : INFINITY;
    }
    __syncthreads();

    for (int r0 = row_beg + warp * VQ_R; r0 < row_end; r0 += VQ_WARPS * VQ_R) {
      float zl[VQ_R][DPL];
#pragma unroll
      for (int r = 0; r < VQ_R; ++r) {
        const int row = min(r0 + r, row_end - 1);       // tail rows are recomputed, never stored twice
        const float* zr = z + (int64_t)row * ld_z + h * DIM;
#pragma unroll
        for (int j = 0; j < DPL; ++j) zl[r][j] = zr[lane + 32 * j];
      }
      float dot[VQ_R][KQ * VEC];
      float zz[VQ_R];
#pragma unroll
      for (int r = 0; r < VQ_R; ++r) {
        zz[r] = 0.f;
#pragma unroll
        for (int c = 0; c < KQ * VEC; ++c) dot[r][c] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
#pragma unroll 4
        for (int dl = 0; dl < 32; ++dl) {
          // sequential over d = 32*j + dl, identical order to oracle/vq_oracle.c
          float ev[KQ * VEC];
          const float* cr = cb + (size_t)(32 * j + dl) * KP + lane * VEC;
#pragma unroll
          for (int q = 0; q < KQ; ++q) {
            if (VEC == 4) {
              const float4 t = *reinterpret_cast<const float4*>(cr + q * 32 * VEC);
              ev[q * VEC + 0] = t.x; ev[q * VEC + 1] = t.y; ev[q * VEC + 2] = t.z; ev[q * VEC + 3] = t.w;
            } else if (VEC == 2) {
              const float2 t = *reinterpret_cast<const float2*>(cr + q * 32 * VEC);
              ev[q * VEC + 0] = t.x; ev[q * VEC + 1] = t.y;
            } else {
              ev[q * VEC] = cr[q * 32 * VEC];
            }
          }
#pragma unroll
          for (int r = 0; r < VQ_R; ++r) {
            const float zd = __shfl_sync(0xffffffffu, zl[r][j], dl);
            zz[r] = fmaf(zd, zd, zz[r]);
#pragma unroll
            for (int c = 0; c < KQ * VEC; ++c) dot[r][c] = fmaf(zd, ev[c], dot[r][c]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < VQ_R; ++r) {
        const int row = r0 + r;
        float best = INFINITY;
        int best_k = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < KQ; ++q)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int k = q * 32 * VEC + lane * VEC + v;
            const float dist = (zz[r] - 2.f * dot[r][q * VEC + v]) + ee[k];
            if (dist < best || (dist == best && k < best_k)) { best = dist; best_k = k; }
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
          if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
        }
        // a row holding NaN / Inf makes every distance NaN and no candidate wins: the reference's (-dist).max(1)
        // then returns index 0 (first NaN); without this the sentinel would index far outside the codebook
        if ((unsigned)best_k >= (unsigned)KP) best_k = 0;
        if (row < row_end) {
          if (lane == 0) idx[(int64_t)row * n_heads + h] = (int64_t)best_k;
#pragma unroll
          for (int j = 0; j < DPL; ++j) {
            const int d = lane + 32 * j;
            const float q = cb[d * KP + best_k];
            const float x = zl[r][j];
            const int64_t o = (int64_t)row * (n_heads * DIM) + h * DIM + d;
            quant_raw[o] = q;
            quant_st[o] = x + (q - x);
            const float dq = q - x;
            const float dv = __fmul_rn(dq, dq);   // no fma contraction with the head sum below (matches the C oracle)
            float* dp = diff + (int64_t)row * DIM + d;
            // sum over heads in head order (python `sum(diffs)`), then / n_heads
            float acc = (h == 0) ? dv : __fadd_rn(*dp, dv);
            if (h == n_heads - 1) acc *= inv_heads;
            *dp = acc;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Search kernel, cluster variant (dim 64, K == padded width: K in {64, 128, 256}; n_heads <= 8).
//
// The generic kernel above walks the heads one after the other inside every CTA and, at the training shape
// (3840 / 960 rows, 4 heads, K = 256), spends most of its time copying four 64 KB codebooks into shared memory with
// scalar loads for 32 rows of work.  Here
//   * grid = (row blocks, heads) with the CTAs of one row block forming a THREAD-BLOCK CLUSTER (cluster dim y =
//     n_heads): the heads of a row block are scored concurrently on neighbouring SMs;
//   * each CTA fetches its ONE codebook with a single cp.async.bulk (TMA unit; completion = mbarrier transaction
//     count) and keeps it for all of its row passes;
//   * the commitment term diff = mean_h (q_h - z_h)^2 needs the heads summed in head order (python `sum(diffs)`):
//     every CTA leaves its per-row squares in shared memory, the cluster synchronises, and each CTA combines a slice
//     of the rows by reading the other CTAs' tiles through DISTRIBUTED SHARED MEMORY in head order -- no global
//     read-modify-write between heads, and the fp32 sum is bit-identical to the sequential one.
// R rows per warp: 4 at small N (more CTAs), 8 at large N (one LDS.128 feeds 32 FMAs); 2 CTAs of 8 warps per SM.
// Distances, their fma order and the tie rule are IDENTICAL to the generic kernel and to oracle/vq_oracle.c.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vq_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int VEC, int KQ, int R, int WARPS, int G>
__global__ void __launch_bounds__(WARPS * 32, 2)
vq_search_cluster_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ embed,
                         float* __restrict__ quant_raw, float* __restrict__ quant_st, float* __restrict__ diff,
                         int64_t* __restrict__ idx, int n_rows, int rows_per_cta) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int DIM = 64;
  constexpr int KP = 32 * VEC * KQ;      // == K
  constexpr int RPP = WARPS * R;         // rows per pass
  constexpr uint32_t CB_BYTES = (uint32_t)DIM * KP * sizeof(float);
  constexpr int DPL = DIM / 32;
  extern __shared__ __align__(128) float smem[];
  float* cb = smem;                                   // [DIM][KP]
  float* ee = cb + (size_t)DIM * KP;                  // [KP]
  float* dvs = ee + KP;                               // [2][G * RPP][DIM]   per-row (q - z)^2 of THIS head
  uint64_t* bar = reinterpret_cast<uint64_t*>(dvs + 2 * G * RPP * DIM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_heads = gridDim.y;
  const int h = blockIdx.y;                           // == rank in the cluster (cluster = all heads of a row block)
  const int row_beg = blockIdx.x * rows_per_cta;
  const int row_end = min(n_rows, row_beg + rows_per_cta);
  const float inv_heads = 1.f / (float)n_heads;

  if (threadIdx.x == 0) {
    const uint32_t b = vq_smem_u32(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CB_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     vq_smem_u32(cb)),
                 "l"(embed + (size_t)h * DIM * KP), "r"(CB_BYTES), "r"(b)
                 : "memory");
  }
  __syncthreads();
  {
    const uint32_t b = vq_smem_u32(bar);
    uint32_t ok;
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(b), "r"(0u)
          : "memory");
    } while (!ok);
  }
  for (int k = threadIdx.x; k < KP; k += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int d = 0; d < DIM; ++d) { float v = cb[d * KP + k]; s = fmaf(v, v, s); }
    ee[k] = s;
  }
  __syncthreads();

  // passes are grouped G at a time: one cluster barrier + one combine per group (a barrier costs about as much as
  // scoring 32 rows); the group buffers are double-buffered so that one barrier per group suffices
  int pass = 0;
  for (int p0 = row_beg; p0 < row_end; p0 += RPP, ++pass) {
    const int grp = pass / G, pig = pass - grp * G;          // group index, pass inside the group
    float* dvg = dvs + (size_t)(grp & 1) * G * RPP * DIM;    // this group's buffer
    float* dv = dvg + (size_t)pig * RPP * DIM;
    const int r0 = p0 + warp * R;
    if (r0 < row_end) {
      float zl[R][DPL];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = min(r0 + r, row_end - 1);       // tail rows are recomputed, never stored
        const float* zr = z + (int64_t)row * ld_z + h * DIM;
#pragma unroll
        for (int j = 0; j < DPL; ++j) zl[r][j] = zr[lane + 32 * j];
      }
      float dot[R][KQ * VEC];
      float zz[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        zz[r] = 0.f;
#pragma unroll
        for (int c = 0; c < KQ * VEC; ++c) dot[r][c] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
#pragma unroll 4
        for (int dl = 0; dl < 32; ++dl) {
          // sequential over d = 32*j + dl, identical order to oracle/vq_oracle.c
          float ev[KQ * VEC];
          const float* cr = cb + (size_t)(32 * j + dl) * KP + lane * VEC;
#pragma unroll
          for (int q = 0; q < KQ; ++q) {
            if (VEC == 4) {
              const float4 t = *reinterpret_cast<const float4*>(cr + q * 32 * VEC);
              ev[q * VEC + 0] = t.x; ev[q * VEC + 1] = t.y; ev[q * VEC + 2] = t.z; ev[q * VEC + 3] = t.w;
            } else {
              const float2 t = *reinterpret_cast<const float2*>(cr + q * 32 * VEC);
              ev[q * VEC + 0] = t.x; ev[q * VEC + 1] = t.y;
            }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float zd = __shfl_sync(0xffffffffu, zl[r][j], dl);
            zz[r] = fmaf(zd, zd, zz[r]);
#pragma unroll
            for (int c = 0; c < KQ * VEC; ++c) dot[r][c] = fmaf(zd, ev[c], dot[r][c]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = r0 + r;
        float best = INFINITY;
        int best_k = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < KQ; ++q)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int k = q * 32 * VEC + lane * VEC + v;
            const float dist = (zz[r] - 2.f * dot[r][q * VEC + v]) + ee[k];
            if (dist < best || (dist == best && k < best_k)) { best = dist; best_k = k; }
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
          if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
        }
        // a row holding NaN / Inf makes every distance NaN and no candidate wins: the reference's (-dist).max(1)
        // then returns index 0 (first NaN); without this the sentinel would index far outside the codebook
        if ((unsigned)best_k >= (unsigned)KP) best_k = 0;
        if (row < row_end) {
          if (lane == 0) idx[(int64_t)row * n_heads + h] = (int64_t)best_k;
#pragma unroll
          for (int j = 0; j < DPL; ++j) {
            const int d = lane + 32 * j;
            const float q = cb[d * KP + best_k];
            const float x = zl[r][j];
            const int64_t o = (int64_t)row * (n_heads * DIM) + h * DIM + d;
            quant_raw[o] = q;
            quant_st[o] = x + (q - x);
            const float dq = q - x;
            dv[(warp * R + r) * DIM + d] = __fmul_rn(dq, dq);   // no fma contraction with the head sum
          }
        }
      }
    }
    const bool group_end = (pig == G - 1) || (p0 + RPP >= row_end);
    if (group_end) {
      // every head's squares of this group are in place (release / acquire across the cluster)
      cluster.sync();
      // combine: this CTA owns the rows ri with ri % n_heads == h; heads are added in head order, then / n_heads
      const int g0 = p0 - pig * RPP;                        // first row of the group
      const int rows_here = min((pig + 1) * RPP, row_end - g0);
      for (int e = threadIdx.x; e < (pig + 1) * RPP * DIM; e += blockDim.x) {
        const int ri = e / DIM, d = e - ri * DIM;
        if (ri < rows_here && (ri % n_heads) == h) {
          float acc = *cluster.map_shared_rank(dvg + ri * DIM + d, 0);
          for (int hh = 1; hh < n_heads; ++hh)
            acc = __fadd_rn(acc, *cluster.map_shared_rank(dvg + ri * DIM + d, hh));
          diff[(int64_t)(g0 + ri) * DIM + d] = acc * inv_heads;
        }
      }
      // buffer (grp & 1) is rewritten in group grp + 2, after the cluster.sync of group grp + 1, which every CTA
      // reaches only after finishing the reads above
    }
  }
  cluster.sync();     // nobody exits while a neighbour may still read its shared memory
}

// ------------------------------------------------------------------------------------------------------------
// Search kernel, large-N variant (n_rows >= 128 per SM; K == padded width).  Heads are walked inside the CTA like in
// the generic kernel -- with thousands of rows per CTA the warps never meet at a barrier inside a head, which hides
// the row loads and the epilogue better than the per-pass cluster barrier of the variant above (measured: 498 vs
// 344 GB/s at 2^20 rows) -- but the
// codebook of head h+1 is fetched by ONE cp.async.bulk (TMA unit, no registers, no issue slots) into the second half
// of a double buffer while head h is being scored; completion is an mbarrier transaction count.  R rows per warp:
// 4 at small N (more CTAs), 8 at large N (one LDS.128 feeds 32 FMAs, the loop becomes FP32-FMA bound).
// Arithmetic and its order are IDENTICAL to the generic kernel (and to oracle/vq_oracle.c): indices stay bit-exact.
// ------------------------------------------------------------------------------------------------------------
template <int DIM, int VEC, int KQ, int R>
__global__ void __launch_bounds__(VQ_WARPS * 32)
vq_search_bulk_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ embed,
                      float* __restrict__ quant_raw, float* __restrict__ quant_st, float* __restrict__ diff,
                      int64_t* __restrict__ idx, int n_rows, int n_heads, int rows_per_cta) {
  extern __shared__ __align__(128) float smem[];
  constexpr int KP = 32 * VEC * KQ;      // == K
  constexpr uint32_t CB_BYTES = (uint32_t)DIM * KP * sizeof(float);
  float* ee = smem + 2 * (size_t)DIM * KP;                 // [KP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ee + KP);   // [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_beg = blockIdx.x * rows_per_cta;
  const int row_end = min(n_rows, row_beg + rows_per_cta);
  const float inv_heads = 1.f / (float)n_heads;
  constexpr int DPL = DIM / 32;  // dims per lane

  auto fetch = [&](int h) {      // one elected thread: whole codebook of head h -> buffer h & 1
    const uint32_t bar = vq_smem_u32(&bars[h & 1]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(CB_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     vq_smem_u32(smem + (size_t)(h & 1) * DIM * KP)),
                 "l"(embed + (size_t)h * DIM * KP), "r"(CB_BYTES), "r"(bar)
                 : "memory");
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(vq_smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(vq_smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fetch(0);
  }
  __syncthreads();

  for (int h = 0; h < n_heads; ++h) {
    const float* cb = smem + (size_t)(h & 1) * DIM * KP;   // [DIM][KP]
    // the other buffer was last read during head h-1, which ended with a __syncthreads
    if (threadIdx.x == 0 && h + 1 < n_heads) fetch(h + 1);
    {
      const uint32_t bar = vq_smem_u32(&bars[h & 1]), parity = (uint32_t)(h >> 1) & 1u;
      uint32_t ok;
      do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
      } while (!ok);
    }
    for (int k = threadIdx.x; k < KP; k += blockDim.x) {
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < DIM; ++d) { float v = cb[d * KP + k]; s = fmaf(v, v, s); }
      ee[k] = s;
    }
    __syncthreads();

    for (int r0 = row_beg + warp * R; r0 < row_end; r0 += VQ_WARPS * R) {
      float zl[R][DPL];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = min(r0 + r, row_end - 1);       // tail rows are recomputed, never stored twice
        const float* zr = z + (int64_t)row * ld_z + h * DIM;
#pragma unroll
        for (int j = 0; j < DPL; ++j) zl[r][j] = zr[lane + 32 * j];
      }
      float dot[R][KQ * VEC];
      float zz[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        zz[r] = 0.f;
#pragma unroll
        for (int c = 0; c < KQ * VEC; ++c) dot[r][c] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
#pragma unroll 4
        for (int dl = 0; dl < 32; ++dl) {
          // sequential over d = 32*j + dl, identical order to oracle/vq_oracle.c
          float ev[KQ * VEC];
          const float* cr = cb + (size_t)(32 * j + dl) * KP + lane * VEC;
#pragma unroll
          for (int q = 0; q < KQ; ++q) {
            if (VEC == 4) {
              const float4 t = *reinterpret_cast<const float4*>(cr + q * 32 * VEC);
              ev[q * VEC + 0] = t.x; ev[q * VEC + 1] = t.y; ev[q * VEC + 2] = t.z; ev[q * VEC + 3] = t.w;
            } else if (VEC == 2) {
              const float2 t = *reinterpret_cast<const float2*>(cr + q * 32 * VEC);
              ev[q * VEC + 0] = t.x; ev[q * VEC + 1] = t.y;
            } else {
              ev[q * VEC] = cr[q * 32 * VEC];
            }
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float zd = __shfl_sync(0xffffffffu, zl[r][j], dl);
            zz[r] = fmaf(zd, zd, zz[r]);
#pragma unroll
            for (int c = 0; c < KQ * VEC; ++c) dot[r][c] = fmaf(zd, ev[c], dot[r][c]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = r0 + r;
        float best = INFINITY;
        int best_k = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < KQ; ++q)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int k = q * 32 * VEC + lane * VEC + v;
            const float dist = (zz[r] - 2.f * dot[r][q * VEC + v]) + ee[k];
            if (dist < best || (dist == best && k < best_k)) { best = dist; best_k = k; }
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
          if (ob < best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
        }
        // a row holding NaN / Inf makes every distance NaN and no candidate wins: the reference's (-dist).max(1)
        // then returns index 0 (first NaN); without this the sentinel would index far outside the codebook
        if ((unsigned)best_k >= (unsigned)KP) best_k = 0;
        if (row < row_end) {
          if (lane == 0) idx[(int64_t)row * n_heads + h] = (int64_t)best_k;
#pragma unroll
          for (int j = 0; j < DPL; ++j) {
            const int d = lane + 32 * j;
            const float q = cb[d * KP + best_k];
            const float x = zl[r][j];
            const int64_t o = (int64_t)row * (n_heads * DIM) + h * DIM + d;
            quant_raw[o] = q;
            quant_st[o] = x + (q - x);
            const float dq = q - x;
            const float dv = __fmul_rn(dq, dq);   // no fma contraction with the head sum below (matches the C oracle)
            float* dp = diff + (int64_t)row * DIM + d;
            // sum over heads in head order (python `sum(diffs)`), then / n_heads
            float acc = (h == 0) ? dv : __fadd_rn(*dp, dv);
            if (h == n_heads - 1) acc *= inv_heads;
            *dp = acc;
          }
        }
      }
    }
    __syncthreads();   // every read of this head's codebook and norms is done
  }
}

// One CTA per (head, codeword): masked count and masked sum of the rows assigned to it, then the EMA.
template <int DIM>
__global__ void __launch_bounds__(256)
vq_ema_accum_kernel(const float* __restrict__ z, int64_t ld_z, const int64_t* __restrict__ idx,
                    const int* __restrict__ lengths, int batch, int t, int n_heads, int K, float decay,
                    float* __restrict__ cluster_size, float* __restrict__ embed_avg) {
  const int h = blockIdx.x / K, k = blockIdx.x % K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  constexpr int DPL = DIM / 32;
  __shared__ float s_sum[8][DIM];
  __shared__ float s_cnt[8];
  float acc[DPL];
#pragma unroll
  for (int j = 0; j < DPL; ++j) acc[j] = 0.f;
  float cnt = 0.f;
  const int n_rows = batch * t;
  // each warp scans a contiguous slab so the summation order is fixed by (warp, row)
  const int slab = ((n_rows + nwarps - 1) / nwarps + 31) / 32 * 32;
  const int rbeg = warp * slab, rend = min(n_rows, rbeg + slab);
  for (int r0 = rbeg; r0 < rend; r0 += 32) {
    const int r = r0 + lane;
    bool hit = false;
    if (r < rend) {
      const int b = r / t, i = r - b * t;
      hit = (i < lengths[b]) && (idx[(int64_t)r * n_heads + h] == (int64_t)k);
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    while (m) {
      // up to four hit rows per round trip (synthetic / early-training batches put hundreds of rows on one
      // codeword: one dependent load per hit made this kernel latency-bound); accumulated in row order
      int sr[4];
      float v[4][DPL];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sr[u] = m ? __ffs(m) - 1 : -1;
        m &= m - 1;            // (0 stays 0)
        if (sr[u] >= 0) {
          const float* zr = z + (int64_t)(r0 + sr[u]) * ld_z + h * DIM;
#pragma unroll
          for (int j = 0; j < DPL; ++j) v[u][j] = zr[lane + 32 * j];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (sr[u] >= 0) {
#pragma unroll
          for (int j = 0; j < DPL; ++j) acc[j] += v[u][j];
          cnt += 1.f;
        }
    }
  }
#pragma unroll
  for (int j = 0; j < DPL; ++j) s_sum[warp][lane + 32 * j] = acc[j];
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x < DIM) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += s_sum[w][threadIdx.x];
    float* ea = embed_avg + ((size_t)h * DIM + threadIdx.x) * K + k;
    *ea = *ea * decay + s * (1.f - decay);   // mul_(decay).add_(sum, alpha=1-decay)
  }
  if (threadIdx.x == 0) {
    float c = 0.f;
    for (int w = 0; w < nwarps; ++w) c += s_cnt[w];
    float* cs = cluster_size + (size_t)h * K + k;
    *cs = *cs * decay + c * (1.f - decay);
  }
}

// One CTA per head: Laplace-smoothed renormalisation and codebook overwrite (modules.py:49-57)
template <int DIM>
__global__ void __launch_bounds__(256)
vq_ema_finalize_kernel(const float* __restrict__ cluster_size, const float* __restrict__ embed_avg,
                       float* __restrict__ embed, int K, float eps) {
  const int h = blockIdx.x;       // gridDim.y CTAs share a head (each recomputes n: K loads)
  __shared__ float s_n;
  if (threadIdx.x < 32) {
    // fixed-order sum: lane-strided partials, then a shuffle tree
    float s = 0.f;
    for (int k = threadIdx.x; k < K; k += 32) s += cluster_size[(size_t)h * K + k];
    s = warp_sum(s);
    if (threadIdx.x == 0) s_n = s;
  }
  __syncthreads();
  const float n = s_n;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < DIM * K; e += gridDim.y * blockDim.x) {
    const int k = e % K;
    const float cs = (cluster_size[(size_t)h * K + k] + eps) / (n + K * eps) * n;
    embed[(size_t)h * DIM * K + e] = embed_avg[(size_t)h * DIM * K + e] / cs;
  }
}

__global__ void vq_backward_kernel(const float* __restrict__ g_quant, const float* __restrict__ g_diff,
                                   const float* __restrict__ z, const float* __restrict__ q,
                                   float* __restrict__ gz, int64_t total, int n_heads, int dim) {
  const float c = 2.f / (float)n_heads;
  const int width = n_heads * dim;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / width;
    const int d = (int)(e % width) % dim;
    float g = g_quant ? g_quant[e] : 0.f;
    if (g_diff) g += c * g_diff[r * dim + d] * (z[e] - q[e]);
    gz[e] = g;
  }
}

// triplet loss per (row, head); one warp per (row, head).  loss(row) = mean_h  red_k  mask * clamp(pos - dist_k + margin, 0) / dim
template <int DIM>
__global__ void __launch_bounds__(256)
vq_triple_kernel(const float* __restrict__ pred, int64_t ld_pred, const float* __restrict__ embed,
                 const int64_t* __restrict__ target, float* __restrict__ loss_rh, float* __restrict__ gpred,
                 int n_rows, int n_heads, int K, float margin, int reduce_mean) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_rows * n_heads) return;
  const int r = gw / n_heads, h = gw % n_heads;
  constexpr int DPL = DIM / 32;
  const float* pr = pred + (int64_t)r * ld_pred + h * DIM;
  const float* e_h = embed + (size_t)h * DIM * K;
  const int tk = (int)target[(int64_t)r * n_heads + h];
  float pl[DPL], tl[DPL];
  float zz = 0.f, pos = 0.f;
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    pl[j] = pr[lane + 32 * j];
    tl[j] = e_h[(size_t)(lane + 32 * j) * K + tk];
    zz = fmaf(pl[j], pl[j], zz);
    const float dd = pl[j] - tl[j];
    pos = fmaf(dd, dd, pos);
  }
  zz = warp_sum(zz);
  pos = warp_sum(pos);
  const float scale = (reduce_mean ? 1.f / (float)K : 1.f) / (float)DIM / (float)n_heads;
  float lsum = 0.f;
  float gacc[DPL];
#pragma unroll
  for (int j = 0; j < DPL; ++j) gacc[j] = 0.f;
  float nactive = 0.f;  // number of active hinge terms (for d pos / d pred)
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    float dot = 0.f, eek = 0.f;
    for (int d = 0; d < DIM; ++d) {
      const float pd = __shfl_sync(0xffffffffu, pl[d >> 5], d & 31);
      if (k < K) {
        const float ev = e_h[(size_t)d * K + k];
        dot = fmaf(pd, ev, dot);
        eek = fmaf(ev, ev, eek);
      }
    }
    bool active = false;
    if (k < K) {
      const float dist = (zz - 2.f * dot) + eek;
      const float tl_ = pos - dist;
      if (tl_ != 0.f) {
        const float v = tl_ + margin;
        if (v > 0.f) { lsum += v; active = true; }
      }
    }
    if (gpred) {
      // d(-dist_k)/d pred = -2 (pred - e_k)
      unsigned m = __ballot_sync(0xffffffffu, active);
      nactive += (float)__popc(m);
      while (m) {
        const int s = __ffs(m) - 1;
        m &= m - 1;
#pragma unroll
        for (int j = 0; j < DPL; ++j)
          gacc[j] -= 2.f * (pl[j] - e_h[(size_t)(lane + 32 * j) * K + (k0 + s)]);
      }
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0) loss_rh[gw] = lsum * scale;
  if (gpred) {
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
      const float g = gacc[j] + nactive * 2.f * (pl[j] - tl[j]);
      gpred[(int64_t)r * (n_heads * DIM) + h * DIM + lane + 32 * j] = g * scale;
    }
  }
}

}  // namespace
}  // namespace msmc

using namespace msmc;

extern "C" int msmc_vq_search(const float* z, int64_t ld_z, const float* embed, float* quant_raw,
                              float* quant_st, float* diff, int64_t* idx, int32_t n_rows, int32_t n_heads,
                              int32_t dim, int32_t n_embed, void* stream) {
  MSMC_REQUIRE(z && embed && quant_raw && quant_st && diff && idx);
  MSMC_REQUIRE(n_rows > 0 && n_heads > 0 && n_embed > 0 && n_embed <= 512);
  if (dim != 64 && dim != 32 && dim != 128 && dim != 256) return MSMC_ERR_UNSUPPORTED;
  // padded codebook width Kp = 32 * VEC * KQ
  int vec, kq;
  if (n_embed <= 32) { vec = 1; kq = 1; }
  else if (n_embed <= 64) { vec = 2; kq = 1; }
  else if (n_embed <= 128) { vec = 4; kq = 1; }
  else if (n_embed <= 256) { vec = 4; kq = 2; }
  else { vec = 4; kq = 4; }
  const int kp = 32 * vec * kq;
  cudaStream_t st = (cudaStream_t)stream;
  static const int use_cluster = [] { const char* e = getenv("MSMC_VQ_CLUSTER"); return e ? atoi(e) : 1; }();
  if (use_cluster && dim == 64 && n_embed == kp && vec >= 2 && kp <= 256 && n_heads <= 8 &&
      (reinterpret_cast<uintptr_t>(embed) & 15) == 0) {
    // cluster variant (the training configurations: K = 64 / 128 / 256 per head, dim 64, 1-8 heads)
    const int sms = num_sms();
    if (n_rows >= sms * 128) {
      // large N: heads inside the CTA, double-buffered bulk-staged codebooks, 8 rows per warp, one CTA per SM
      const size_t bsmem = (2 * (size_t)dim * kp + kp) * sizeof(float) + 2 * sizeof(uint64_t);
      const int quantum = VQ_WARPS * 8;
      int rows_per_cta = std::max(quantum, (int)ceil_div(n_rows, sms));
      rows_per_cta = ceil_div(rows_per_cta, quantum) * quantum;
      const int grid = ceil_div(n_rows, rows_per_cta);
#define LAUNCH_VQB(V, Q)                                                                                       \
  do {                                                                                                        \
    cudaFuncSetAttribute(vq_search_bulk_kernel<64, V, Q, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                         (int)bsmem);                                                                         \
    vq_search_bulk_kernel<64, V, Q, 8><<<grid, VQ_WARPS * 32, bsmem, st>>>(z, ld_z, embed, quant_raw, quant_st, \
                                                                           diff, idx, n_rows, n_heads,        \
                                                                           rows_per_cta);                     \
  } while (0)
      if (vec == 2) LAUNCH_VQB(2, 1);
      else if (kq == 1) LAUNCH_VQB(4, 1);
      else LAUNCH_VQB(4, 2);
#undef LAUNCH_VQB
      MSMC_CHECK_LAUNCH();
      return MSMC_OK;
    }
    // 8 rows per warp halve the shared-memory traffic per FMA (at 4 rows the LDS.128 stream takes as long as the
    // FMAs); 4 rows keep enough CTAs in flight when there are fewer than ~2k rows
    const bool r8 = n_rows >= 2048;
    const int warps = 8, rpw = r8 ? 8 : 4, grp = r8 ? 1 : 2;
    const int rpp = warps * rpw;
    const size_t csmem = ((size_t)dim * kp + kp + 2 * (size_t)grp * rpp * dim) * sizeof(float) + sizeof(uint64_t);
    // one wave of clusters: 2 CTAs per SM (<= 97 KB of shared memory, <= 128 registers x 256 threads each)
    const int max_clusters = std::max(1, 2 * sms / n_heads);
    int gx = std::min(ceil_div(n_rows, rpp), max_clusters);
    int rows_per_cta = ceil_div(ceil_div(n_rows, gx), rpp) * rpp;
    gx = ceil_div(n_rows, rows_per_cta);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)gx, (unsigned)n_heads, 1);
    cfg.blockDim = dim3((unsigned)(warps * 32), 1, 1);
    cfg.dynamicSmemBytes = csmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)n_heads;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define LAUNCH_VQC(V, Q, R_, W_, G_)                                                                             \
  do {                                                                                                        \
    cudaFuncSetAttribute(vq_search_cluster_kernel<V, Q, R_, W_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)csmem);                                                                         \
    if (cudaLaunchKernelEx(&cfg, vq_search_cluster_kernel<V, Q, R_, W_, G_>, z, ld_z, embed, quant_raw, quant_st, \
                           diff, idx, (int)n_rows, rows_per_cta) != cudaSuccess)                              \
      return MSMC_ERR_LAUNCH;                                                                                 \
  } while (0)
    if (vec == 2) { if (r8) LAUNCH_VQC(2, 1, 8, 8, 1); else LAUNCH_VQC(2, 1, 4, 8, 2); }
    else if (kq == 1) { if (r8) LAUNCH_VQC(4, 1, 8, 8, 1); else LAUNCH_VQC(4, 1, 4, 8, 2); }
    else { if (r8) LAUNCH_VQC(4, 2, 8, 8, 1); else LAUNCH_VQC(4, 2, 4, 8, 2); }
#undef LAUNCH_VQC
    MSMC_CHECK_LAUNCH();
    return MSMC_OK;
  }
  const size_t smem = ((size_t)dim * kp + kp) * sizeof(float);
  if (smem > 200 * 1024) return MSMC_ERR_UNSUPPORTED;
  // one CTA per SM where possible; a warp handles VQ_R rows per pass
  const int quantum = VQ_WARPS * VQ_R;
  int rows_per_cta = std::max(quantum, (int)ceil_div(n_rows, num_sms()));
  rows_per_cta = ceil_div(rows_per_cta, quantum) * quantum;
  const int grid = ceil_div(n_rows, rows_per_cta);
#define LAUNCH_VQ3(D, V, Q)                                                                                   \
  do {                                                                                                        \
    if (smem > 48 * 1024)                                                                                     \
      cudaFuncSetAttribute(vq_search_kernel<D, V, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    vq_search_kernel<D, V, Q><<<grid, VQ_WARPS * 32, smem, st>>>(z, ld_z, embed, quant_raw, quant_st, diff,  \
                                                                  idx, n_rows, n_heads, n_embed, rows_per_cta);\
  } while (0)
#define LAUNCH_VQ(D)                                    \
  do {                                                  \
    if (vec == 1) LAUNCH_VQ3(D, 1, 1);                  \
    else if (vec == 2) LAUNCH_VQ3(D, 2, 1);             \
    else if (kq == 1) LAUNCH_VQ3(D, 4, 1);              \
    else if (kq == 2) LAUNCH_VQ3(D, 4, 2);              \
    else LAUNCH_VQ3(D, 4, 4);                           \
  } while (0)
  switch (dim) {
    case 32: LAUNCH_VQ(32); break;
    case 64: LAUNCH_VQ(64); break;
    case 128: LAUNCH_VQ(128); break;
    default: LAUNCH_VQ(256); break;
  }
#undef LAUNCH_VQ
#undef LAUNCH_VQ3
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_vq_ema_update(const float* z, int64_t ld_z, const int64_t* idx, const int32_t* lengths,
                                  int32_t batch, int32_t t, int32_t n_heads, int32_t dim, int32_t n_embed,
                                  float decay, float eps, float* cluster_size, float* embed_avg, float* embed,
                                  void* stream) {
  MSMC_REQUIRE(z && idx && lengths && cluster_size && embed_avg && embed);
  MSMC_REQUIRE(batch > 0 && t > 0 && n_heads > 0 && n_embed > 0);
  if (dim != 64 && dim != 32 && dim != 128 && dim != 256) return MSMC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_EMA(D)                                                                                        \
  do {                                                                                                       \
    vq_ema_accum_kernel<D><<<n_heads * n_embed, 256, 0, st>>>(z, ld_z, idx, lengths, batch, t, n_heads,      \
                                                               n_embed, decay, cluster_size, embed_avg);      \
    vq_ema_finalize_kernel<D><<<dim3(n_heads, 16), 256, 0, st>>>(cluster_size, embed_avg, embed, n_embed, eps);        \
  } while (0)
  switch (dim) {
    case 32: LAUNCH_EMA(32); break;
    case 64: LAUNCH_EMA(64); break;
    case 128: LAUNCH_EMA(128); break;
    default: LAUNCH_EMA(256); break;
  }
#undef LAUNCH_EMA
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_vq_backward(const float* g_quant, const float* g_diff, const float* z,
                                const float* quant_raw, float* gz, int32_t n_rows, int32_t n_heads,
                                int32_t dim, void* stream) {
  MSMC_REQUIRE(z && quant_raw && gz && n_rows > 0 && n_heads > 0 && dim > 0);
  const int64_t total = (int64_t)n_rows * n_heads * dim;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
  vq_backward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g_quant, g_diff, z, quant_raw, gz, total,
                                                               n_heads, dim);
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}

extern "C" int msmc_vq_triple_loss(const float* pred, int64_t ld_pred, const float* embed,
                                   const int64_t* target, float* loss, float* gpred, int32_t n_rows,
                                   int32_t n_heads, int32_t dim, int32_t n_embed, float margin,
                                   int32_t reduce_mean, void* stream) {
  MSMC_REQUIRE(pred && embed && target && loss && n_rows > 0 && n_heads > 0 && n_embed > 0);
  if (dim != 64 && dim != 32 && dim != 128 && dim != 256) return MSMC_ERR_UNSUPPORTED;
  const int warps = n_rows * n_heads;
  const int blocks = ceil_div(warps, 8);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dim) {
    case 32: vq_triple_kernel<32><<<blocks, 256, 0, st>>>(pred, ld_pred, embed, target, loss, gpred, n_rows, n_heads, n_embed, margin, reduce_mean); break;
    case 64: vq_triple_kernel<64><<<blocks, 256, 0, st>>>(pred, ld_pred, embed, target, loss, gpred, n_rows, n_heads, n_embed, margin, reduce_mean); break;
    case 128: vq_triple_kernel<128><<<blocks, 256, 0, st>>>(pred, ld_pred, embed, target, loss, gpred, n_rows, n_heads, n_embed, margin, reduce_mean); break;
    default: vq_triple_kernel<256><<<blocks, 256, 0, st>>>(pred, ld_pred, embed, target, loss, gpred, n_rows, n_heads, n_embed, margin, reduce_mean); break;
  }
  MSMC_CHECK_LAUNCH();
  return MSMC_OK;
}
