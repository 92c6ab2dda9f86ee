/*
 * msmc_b200.h -- C-ABI of the B200-native MSMC-VQ-GAN training hot path.
 *
 * Every entry point is `extern "C"`, takes raw DEVICE pointers, sizes and an opaque
 * `cudaStream_t` (passed as void*), allocates nothing, never synchronises, and returns
 * 0 on success / a non-zero msmc_status on a bad argument or a CUDA launch error
 * (the host side raises on non-zero).  No torch types appear in any signature.
 *
 * The reference (hhguo/MSMC-TTS) has no FFI of its own: its hot path is torch.nn calls.
 * Each entry point therefore cites the reference Python call site(s) it replaces.
 * Layout convention: activations are CHANNELS-LAST, (B, H, W, C) row-major with an
 * explicit row pitch in elements (1-D sequences use H == 1).
 */
#ifndef MSMC_B200_H
#define MSMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MSMC_OK = 0,
  MSMC_ERR_BAD_ARG = 1,
  MSMC_ERR_LAUNCH = 2,
  MSMC_ERR_UNSUPPORTED = 3
} msmc_status;

/* element-wise transforms a conv kernel can apply while loading an operand / storing a result */
typedef enum {
  MSMC_XF_NONE = 0,
  MSMC_XF_LRELU = 1,       /* v -> v>0 ? v : slope*v                       (F.leaky_relu)          */
  MSMC_XF_RELU = 2,        /* v -> max(v,0)                                                        */
  MSMC_XF_TANH = 3,        /* v -> tanh(v)                                                         */
  MSMC_XF_MUL_DLRELU = 4,  /* v -> v * (aux>0 ? 1 : slope)   aux = forward INPUT of the lrelu      */
  MSMC_XF_MUL_DRELU = 5,   /* v -> v * (aux>0)               aux = forward OUTPUT of the relu      */
  MSMC_XF_MUL_DTANH = 6    /* v -> v * (1-aux*aux)           aux = forward OUTPUT of the tanh      */
} msmc_xform;

/*
 * Geometry of one convolution-shaped contraction over channels-last tensors.
 *   src : (B, Hs, Ws, Cs) pitch ld_src      dst : (B, Hd, Wd, Cd) pitch ld_dst
 *   transposed == 0 (conv forward form):   dst(hd,wd) gathers src(hd*sh + kh*dh - ph, wd*sw + kw*dw - pw)
 *   transposed == 1 (conv-transpose form): dst(hd,wd) gathers src((hd + ph - kh*dh)/sh, (wd + pw - kw*dw)/sw)
 *                                          when divisible and in range
 *   weight element for (kh, kw, cs, cd) lives at  w[kh*ws_kh + kw*ws_kw + cs*ws_cs + cd*ws_cd]
 * so torch's native Conv/ConvTranspose/Linear weight layouts are consumed in place.
 */
typedef struct {
  int32_t B, Hs, Ws, Cs;
  int32_t Hd, Wd, Cd;
  int32_t KH, KW;
  int32_t sh, sw, dh, dw, ph, pw;
  int32_t pad_reflect;      /* 0: zeros outside, 1: reflect (ReflectionPad2d semantics); forward form only */
  int32_t transposed;
  int64_t ld_src, ld_dst, ld_res;
  int64_t ld_saux, ld_daux;  /* pitches of the src-shaped / dst-shaped aux tensors */
  int64_t ws_kh, ws_kw, ws_cs, ws_cd;
  int32_t src_xf;  float src_slope;   /* transform of src values on load (aux tensor shaped like src)   */
  int32_t dst_xf;  float dst_slope;   /* transform of the result before the residual add (aux like dst) */
} msmc_conv_geom;

/* ---------------------------------------------------------------------------------------------
 * Convolution family (a5-a12, a14, a15, a19; reference: torch.nn.Conv1d/ConvTranspose1d/
 * Conv2d/Linear calls at hifigan/generator.py:22-36,40-55, hifigan/common.py:21-51,
 * hifigan/discriminator.py:19-69,126-154, acoustic_models/transformer.py:222,233,346-350,
 * vqgantts/msmc_vqgan.py:116-118,132-134,285,307, vqgantts/modules.py:207-221).
 * dst = dst_xf( sum_{taps,cs} src_xf(src) * w  + bias ) + residual
 * ------------------------------------------------------------------------------------------- */
int msmc_conv_forward(const msmc_conv_geom* g, const float* src, const float* src_aux,
                      const float* w, const float* bias, const float* residual,
                      const float* dst_aux, float* dst, void* stream);

/* Tensor-core path of msmc_conv_forward (forward form only, Cs % 32 == 0): tcgen05.mma kind::tf32 with the
 * accumulator in TMEM; `split` != 0 selects 3xTF32 (fp32-accurate), 0 plain TF32.  The weights come as swizzled
 * shared-memory tile images built by msmc_weight_image from the GEMM layout [T][Cs][Cd]:
 *   role 0: forward operand;  role 1: data-gradient operand of a stride-1 conv (taps reversed, channels swapped),
 * so the data gradient is again an msmc_conv_forward_umma call with padding d*(K-1)-p.
 * msmc_umma_tile_n(Cd, rows) proposes the N tile (32/64/128) that fills the SMs; image and kernel must be given
 * the same BN. */
int msmc_umma_tile_n(int32_t out_channels, int64_t rows);
int64_t msmc_weight_image_elems(int32_t T, int32_t Cs, int32_t Cd, int32_t role, int32_t split, int32_t BN);
int msmc_weight_image(const float* w_gemm, float* image, int32_t T, int32_t Cs, int32_t Cd, int32_t role,
                      int32_t split, int32_t BN, void* stream);
/* every operand image of a sub-network in ONE launch.  jobs = DEVICE table of int64 words, 9 per job:
 * w_gemm, image, T, Cs, Cd, BN, role, split, blk0 (blk0 = sum over the preceding jobs of ceil(elements / 1024) with
 * elements = msmc_weight_image_elems(.., split = 0, ..)); total_blocks = that sum over all jobs */
int msmc_weight_image_multi(const int64_t* jobs, int32_t n_jobs, int64_t total_blocks, void* stream);
int msmc_conv_forward_umma(const msmc_conv_geom* g, const float* src, const float* src_aux, const float* wimg,
                           const float* bias, const float* residual, const float* dst_aux, float* dst,
                           int32_t split, int32_t BN, void* stream);

/* stride-1 convolutions whose taps are constant row shifts in flattened pixel space (1-D convs, (k,1) convs) reuse ONE
 * staged operand tile for all taps (shifted shared-memory descriptors); same contract as msmc_conv_forward_umma */
int msmc_conv_reuse_eligible(const msmc_conv_geom* g);
int msmc_conv_forward_umma_reuse(const msmc_conv_geom* g, const float* src, const float* src_aux,
                                 const float* wimg, const float* bias, const float* residual,
                                 const float* dst_aux, float* dst, int32_t split, int32_t BN, void* stream);

/* tensor-core weight gradient (Cs % 32 == 0): both operands are consumed MN-major straight from the channels-last
 * activations; same contract and workspace layout as msmc_conv_wgrad */
int64_t msmc_conv_wgrad_umma_workspace(const msmc_conv_geom* g);
int msmc_conv_wgrad_umma(const msmc_conv_geom* g, const float* src, const float* src_aux, const float* gout,
                         const float* gout_aux, float* dw, float* dbias, float* workspace, int64_t workspace_bytes,
                         int32_t split, void* stream);

/* weight gradient of the forward form: dW[kh,kw,cs,cd] = sum_rows src_xf(src)[gathered] * gout_xf(gout)
 * (gout_xf is given in g->dst_xf with aux = dst_aux).  `workspace` holds split partial sums;
 * query its size with msmc_conv_wgrad_workspace().  dbias (optional) receives column sums of gout. */
int64_t msmc_conv_wgrad_workspace(const msmc_conv_geom* g);
int msmc_conv_wgrad(const msmc_conv_geom* g, const float* src, const float* src_aux,
                    const float* gout, const float* gout_aux, float* dw, float* dbias,
                    float* workspace, int64_t workspace_bytes, void* stream);

/* weight_norm re-parametrisation (torch.nn.utils.weight_norm, dim=0; generator.py:22-36, common.py:24-41,
 * discriminator.py:23-68,127-133, modules.py:207-221):  w[o,:] = g[o] * v[o,:] / ||v[o,:]||  */
/* v is contiguous (O, I, J), the norm runs over (I, J); element (o,i,j) of w is written at o*so + i*si + j*sj so
 * the kernel emits the GEMM layout [tap][cs][cd] directly.  g == NULL: plain re-layout (un-normalised weights). */
int msmc_weight_norm_fwd(const float* v, const float* g, float* w, float* inv_norm, int32_t O, int32_t I,
                         int32_t J, int64_t so, int64_t si, int64_t sj, void* stream);
/* the same for EVERY weight of a sub-network in one launch (the reference re-parametrises each layer inside its own
 * forward: ~270 tiny launches per train step).  jobs = DEVICE table of int64 words, 11 per job:
 * v, g (0 = plain re-layout), w, inv_norm (0 = not wanted), so, si, sj, O, I, J, row0 (= sum of O over the
 * preceding jobs); total_rows = sum of O over all jobs.  Bit-identical to msmc_weight_norm_fwd per job. */
int msmc_weight_norm_fwd_multi(const int64_t* jobs, int32_t n_jobs, int64_t total_rows, void* stream);
int msmc_weight_norm_bwd(const float* dw, int64_t so, int64_t si, int64_t sj, const float* v, const float* g,
                         const float* inv_norm, float* dv, float* dg, int32_t O, int32_t I, int32_t J,
                         void* stream);
/* backward of nn.ReflectionPad2d (discriminator.py:20-66) / reflect-padded STFT framing:
 * fold a (B, H+2ph, W+2pw, C) gradient onto (B, H, W, C) */
int msmc_reflect_pad_fold(const float* gpad, float* gx, int32_t B, int32_t H, int32_t W, int32_t C,
                          int32_t ph, int32_t pw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head VQ (a1-a3; reference vqgantts/modules.py:24-67,137-151).
 * z, quant : (n_rows, n_heads*dim) pitch ld;  embed : n_heads x (dim, n_embed) dim-major (the reference's
 * `embed` buffers stacked);  idx : (n_rows, n_heads) int64;  diff : (n_rows, dim) = mean_h (q_h - z_h)^2.
 * quant_st = z + (q - z)   (the straight-through value the reference returns).
 * ------------------------------------------------------------------------------------------- */
int msmc_vq_search(const float* z, int64_t ld_z, const float* embed, float* quant_raw, float* quant_st,
                   float* diff, int64_t* idx, int32_t n_rows, int32_t n_heads, int32_t dim,
                   int32_t n_embed, void* stream);
/* The same search, results identical by construction, as a two-phase tensor-core kernel (the host side selects it
 * from 2048 rows on; csrc/vq_umma.cu) -- all codewords scored with tcgen05 (3xTF32), the minimum accepted when the
 * runner-up lies outside a provable margin, any other row re-scored with the exact sequential-fma arithmetic.
 * dim = 64, n_embed in {64, 128, 256}, 1 / 2 / 4 / 8 heads, ld_z a multiple of 4, 16-byte aligned pointers;
 * MSMC_ERR_UNSUPPORTED otherwise. */
int msmc_vq_search_umma(const float* z, int64_t ld_z, const float* embed, float* quant_raw, float* quant_st,
                        float* diff, int64_t* idx, int32_t n_rows, int32_t n_heads, int32_t dim, int32_t n_embed,
                        void* stream);
/* EMA codebook update (modules.py:35-57).  row r = b*t + i is valid iff i < lengths[b].
 * Updates cluster_size (n_heads,n_embed), embed_avg and embed (n_heads,dim,n_embed) in place. */
int msmc_vq_ema_update(const float* z, int64_t ld_z, const int64_t* idx, const int32_t* lengths,
                       int32_t batch, int32_t t, int32_t n_heads, int32_t dim, int32_t n_embed,
                       float decay, float eps, float* cluster_size, float* embed_avg, float* embed,
                       void* stream);
/* backward of (quant_st, diff) w.r.t. z:  gz = g_quant + (2/H) * g_diff * (z - q) */
int msmc_vq_backward(const float* g_quant, const float* g_diff, const float* z, const float* quant_raw,
                     float* gz, int32_t n_rows, int32_t n_heads, int32_t dim, void* stream);
/* triplet loss of Quantize.compute_triple_loss (modules.py:86-116), 'sum' or 'mean' over codewords */
int msmc_vq_triple_loss(const float* pred, int64_t ld_pred, const float* embed, const int64_t* target,
                        float* loss, float* gpred, int32_t n_rows, int32_t n_heads, int32_t dim,
                        int32_t n_embed, float margin, int32_t reduce_mean, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FFT-block attention (a8; transformer.py:246-328).  qkv : (B, t, n_head*3*d) as produced by the
 * fused QKV Linear (per head: q | k | v), out : (B, t, n_head*d).  Keys j >= lengths[b] are masked (-inf).
 * Dropout on the probabilities uses a counter-based generator keyed by (*seed, call_salt, b, h, q, k).
 * ------------------------------------------------------------------------------------------- */
int msmc_attention_fwd(const float* qkv, const int32_t* lengths, float* out, float* lse,
                       int32_t B, int32_t t, int32_t n_head, int32_t d, float inv_temperature,
                       float drop_p, const uint64_t* seed, uint64_t call_salt, void* stream);
int msmc_attention_bwd(const float* qkv, const int32_t* lengths, const float* out, const float* lse,
                       const float* gout, float* gqkv, int32_t B, int32_t t, int32_t n_head, int32_t d,
                       float inv_temperature, float drop_p, const uint64_t* seed, uint64_t call_salt,
                       void* stream);

/* residual + dropout + LayerNorm + pad mask (transformer.py:275-283,374-380,201-204):
 *   y = mask(b,i) * LN( dropout(a) + r ) ;  rows = B*t, row (b,i) masked to 0 when i >= lengths[b] */
int msmc_add_layernorm_fwd(const float* a, const float* r, const float* gamma, const float* beta,
                           const int32_t* lengths, float* y, float* xhat, float* rstd,
                           int32_t B, int32_t t, int32_t C, float eps, float drop_p,
                           const uint64_t* seed, uint64_t call_salt, void* stream);
/* bytes of `workspace` msmc_add_layernorm_bwd needs (per-CTA partial dgamma/dbeta rows) */
int64_t msmc_add_layernorm_bwd_workspace(int32_t C);
int msmc_add_layernorm_bwd(const float* gy, const float* xhat, const float* rstd, const float* gamma,
                           const int32_t* lengths, float* ga, float* gr, float* dgamma, float* dbeta,
                           float* workspace, int32_t B, int32_t t, int32_t C, float drop_p,
                           const uint64_t* seed, uint64_t call_salt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Spectral front end of the multi-resolution discriminator and MelLoss (a13, a17;
 * utils/audio.py:398-419,348-376, criterions/stft_loss.py:78-107).  The windowed DFT itself is a
 * msmc_conv_forward with Cs = 1; these are the HBM-bound point-wise stages around it.
 *   spec : (rows, 2*Fp) = [re(0..F) pad | im(0..F) pad], Fp >= F (channel padding keeps the spectrum GEMMs on the
 *   tensor-core path); mag (rows, F) = sqrt(max(re^2+im^2, floor_)) (floor_add=0) or sqrt(re^2+im^2+floor_) (=1)
 * ------------------------------------------------------------------------------------------- */
int msmc_spec_magnitude_fwd(const float* spec, float* mag, int64_t rows, int32_t F, int32_t Fp, float floor_,
                            int32_t floor_add, void* stream);
int msmc_spec_magnitude_bwd(const float* gmag, const float* spec, const float* mag, float* gspec,
                            int64_t rows, int32_t F, int32_t Fp, float floor_, int32_t floor_add, void* stream);
/* 'double' domain (audio.py:414-417): out[...,0] = mel, out[...,1] = clamp((20 log10(mel) - ref + 100)/100, 0, 1)
 * mel : (rows, F) -> out : (rows, F, 2) channels-last */
int msmc_mel_double_fwd(const float* mel, float* out, int64_t n, float ref_db, float min_db, void* stream);
int msmc_mel_double_bwd(const float* gout, const float* mel, float* gmel, int64_t n, float ref_db,
                        float min_db, void* stream);
/* out[i] = xf(v[i], aux[i]) for one msmc_xform (pointers 16-byte aligned).  The conv backward uses it to form the
 * pre-activation gradient g * act'(y) of a fused output activation once (reference: autograd of F.leaky_relu /
 * F.relu / torch.tanh after each conv, e.g. discriminator.py:41-47, transformer.py:361) */
int msmc_xform_apply(const float* v, const float* aux, float* out, int64_t n, int32_t xf, float slope, void* stream);
/* log-compression of MelLoss: out = log(max(x, clip)) and its backward */
int msmc_log_clamp_fwd(const float* x, float* y, int64_t n, float clip, void* stream);
int msmc_log_clamp_bwd(const float* gy, const float* x, float* gx, int64_t n, float clip, void* stream);

/* fused multi-tensor Adam / AdamW step with torch.optim semantics (reference trainers/optimizers/__init__.py:9-30
 * builds torch.optim.Adam / AdamW per child module) with the global gradient-norm clip of the trainer folded in
 * (reference trainers/msmctts_trainer.py:203-207: clip_grad_norm_ followed by optimizer.step()).
 * table = device array [4][n_tensors] of pointers (param, grad, exp_avg, exp_avg_sq), sizes[n_tensors] element
 * counts, one CTA per (chunk_tensor[c], chunk_index[c]) chunk of msmc_adam_chunk_elems() elements.
 * steps = DEVICE vector of per-parameter step counters, step_index[t] = slot of tensor t: the call advances the
 * counter of every listed tensor by one (torch advances a parameter's step only when it has a gradient) and uses the
 * advanced value for the bias correction.  lr is a DEVICE scalar.
 * max_norm > 0 (needs partial = n_chunks floats of workspace): gradients are scaled in place by
 * min(1, max_norm / (total_norm + 1e-6)) before the update, total_norm = l2 norm over all listed gradients, also
 * written to norm_out[0] when non-null (torch.nn.utils.clip_grad_norm_ semantics).  Two launches. */
int msmc_adam_chunk_elems(void);
int msmc_adam_multi(const uint64_t* table, int32_t n_tensors, const int64_t* sizes, const int32_t* chunk_tensor,
                    const int32_t* chunk_index, int32_t n_chunks, const int32_t* step_index, float* steps,
                    float* partial, float max_norm, float* norm_out, const float* lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t decoupled, void* stream);
/* feature-matching loss over a list of tensor pairs (reference trainers/msmctts_trainer.py:186-190: the sum of 55
 * F.l1_loss(fake_fmap, real_fmap) terms): out[0] = sum_t mean|a_t - b_t|.  table = device array [2][n_tensors] of
 * pointers (a, b) ([3][n_tensors] with the gradient buffers ga for _bwd), chunks of msmc_l1_chunk_elems() elements,
 * partial = n_chunks floats of workspace; the reduction order is fixed (deterministic).
 * _bwd: ga_t = gout[0] * sign(a_t - b_t) / n_t   (gout is a DEVICE scalar) */
int msmc_l1_chunk_elems(void);
int msmc_l1_multi_fwd(const uint64_t* table, int32_t n_tensors, const int64_t* sizes, const int32_t* chunk_tensor,
                      const int32_t* chunk_index, int32_t n_chunks, float* partial, float* out, void* stream);
int msmc_l1_multi_bwd(const uint64_t* table, int32_t n_tensors, const int64_t* sizes, const int32_t* chunk_tensor,
                      const int32_t* chunk_index, int32_t n_chunks, const float* gout, void* stream);
/* forward STFT framing (torch.stft center / reflect padding, audio.py:399, stft_loss.py:88-99): gather the overlapping
 * frames of x (B, L) into a dense (B*frames, win_p) matrix (columns >= win are zero) so the windowed DFT is one GEMM */
int msmc_frame_unfold(const float* x, float* frames_out, int32_t B, int32_t L, int32_t frames, int32_t win,
                      int32_t win_p, int32_t hop, int32_t pad, void* stream);
/* backward of reflect-padded STFT framing (torch.stft center/reflect, audio.py:399; stft_loss.py:88-99): overlap-add
 * the per-frame time-domain gradients gframes (B, frames, win) onto x (B, L) and fold the reflected borders */
int msmc_overlap_add_fold(const float* gframes, float* gx, int32_t B, int32_t frames, int32_t win, int32_t hop,
                          int32_t L, int32_t pad, void* stream);

/* gated activation of the WaveNet-style ResStack (modules.py:172-179): out = tanh(x[:, :C]) * sigmoid(x[:, C:]) */
int msmc_gated_act_fwd(const float* x, float* y, int64_t rows, int32_t C, void* stream);
int msmc_gated_act_bwd(const float* gy, const float* x, float* gx, int64_t rows, int32_t C, void* stream);

/* library / device info */
int msmc_version(void);
int msmc_num_sms(void);

#ifdef __cplusplus
}
#endif
#endif /* MSMC_B200_H */
