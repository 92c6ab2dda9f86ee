/*
 * vq_oracle.c -- CPU restatement (plain C) of the reference's VQ quantiser.  TEST INFRASTRUCTURE ONLY:
 * imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product path.
 *
 * Follows hhguo/MSMC-TTS msmctts/networks/vqgantts/modules.py:
 *   search      : lines 25-33   dist = ||z||^2 - 2 z.E + ||E||^2 ; ind = argmax(-dist) ; quantize = E[:, ind]
 *   EMA update  : lines 35-57   masked one-hot counts / sums, decay, Laplace smoothing, overwrite embed
 *   outputs     : lines 59-60   diff = (q - z)^2 ; quantize = z + (q - z)
 *   multi-head  : lines 137-151 chunk last dim, per-head quantiser, diff = sum(diffs)/H, idx stacked
 *
 * Distances are evaluated in fp32 with a fixed sequential fma order (d = 0..dim-1); the CUDA kernel
 * (csrc/vq.cu) uses the same order so code indices agree bit for bit.  The reference's own order is whatever
 * its BLAS picks; oracle/make_golden.py pins this file against the reference's outputs (indices equal on every
 * golden vector; smallest top-2 distance gap recorded in the fixture).
 *
 * embed layout: n_heads x (dim, n_embed), dim-major -- the reference's `embed` buffers stacked.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void vq_oracle_search(const float* z, int64_t ld_z, const float* embed, float* quant_raw, float* quant_st,
                      float* diff, int64_t* idx, int n_rows, int n_heads, int dim, int n_embed) {
  float* ee = (float*)malloc(sizeof(float) * (size_t)n_embed);
  for (int h = 0; h < n_heads; ++h) {
    const float* E = embed + (size_t)h * dim * n_embed;
    for (int k = 0; k < n_embed; ++k) {
      float s = 0.f;
      for (int d = 0; d < dim; ++d) s = fmaf(E[(size_t)d * n_embed + k], E[(size_t)d * n_embed + k], s);
      ee[k] = s;
    }
    for (int r = 0; r < n_rows; ++r) {
      const float* zr = z + (size_t)r * ld_z + (size_t)h * dim;
      float zz = 0.f;
      for (int d = 0; d < dim; ++d) zz = fmaf(zr[d], zr[d], zz);
      float best = INFINITY;
      int best_k = 0;
      for (int k = 0; k < n_embed; ++k) {
        float dot = 0.f;
        for (int d = 0; d < dim; ++d) dot = fmaf(zr[d], E[(size_t)d * n_embed + k], dot);
        const float dist = (zz - 2.f * dot) + ee[k];
        if (dist < best) { best = dist; best_k = k; }   /* first minimum wins, like (-dist).max(1) */
      }
      idx[(size_t)r * n_heads + h] = best_k;
      for (int d = 0; d < dim; ++d) {
        const float q = E[(size_t)d * n_embed + best_k];
        const float x = zr[d];
        const size_t o = (size_t)r * n_heads * dim + (size_t)h * dim + d;
        quant_raw[o] = q;
        quant_st[o] = x + (q - x);
        const float dv = (q - x) * (q - x);
        float* dp = diff + (size_t)r * dim + d;
        float acc = (h == 0) ? dv : (*dp + dv);
        if (h == n_heads - 1) acc *= 1.f / (float)n_heads;
        *dp = acc;
      }
    }
  }
  free(ee);
}

/* modules.py:35-57; row r = b*t + i is used iff i < lengths[b] */
void vq_oracle_ema(const float* z, int64_t ld_z, const int64_t* idx, const int32_t* lengths, int batch, int t,
                   int n_heads, int dim, int n_embed, float decay, float eps, float* cluster_size,
                   float* embed_avg, float* embed) {
  double* sum = (double*)malloc(sizeof(double) * (size_t)dim * n_embed);
  double* cnt = (double*)malloc(sizeof(double) * (size_t)n_embed);
  for (int h = 0; h < n_heads; ++h) {
    memset(sum, 0, sizeof(double) * (size_t)dim * n_embed);
    memset(cnt, 0, sizeof(double) * (size_t)n_embed);
    for (int b = 0; b < batch; ++b)
      for (int i = 0; i < t && i < lengths[b]; ++i) {
        const size_t r = (size_t)b * t + i;
        const int k = (int)idx[r * n_heads + h];
        cnt[k] += 1.0;
        for (int d = 0; d < dim; ++d) sum[(size_t)d * n_embed + k] += z[r * ld_z + (size_t)h * dim + d];
      }
    float* cs = cluster_size + (size_t)h * n_embed;
    float* ea = embed_avg + (size_t)h * dim * n_embed;
    float* em = embed + (size_t)h * dim * n_embed;
    float n = 0.f;
    for (int k = 0; k < n_embed; ++k) { cs[k] = cs[k] * decay + (float)cnt[k] * (1.f - decay); }
    for (int k = 0; k < n_embed; ++k) n += cs[k];
    for (int d = 0; d < dim; ++d)
      for (int k = 0; k < n_embed; ++k) {
        const size_t e = (size_t)d * n_embed + k;
        ea[e] = ea[e] * decay + (float)sum[e] * (1.f - decay);
        const float c = (cs[k] + eps) / (n + n_embed * eps) * n;
        em[e] = ea[e] / c;
      }
  }
  free(sum);
  free(cnt);
}
