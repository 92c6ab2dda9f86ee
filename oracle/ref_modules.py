"""Floating-point oracle: torch-CPU fp32 functional restatement of the MSMC-VQ-GAN hot path.
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).

Every function takes the parameters as a flat dict keyed by the REFERENCE's state_dict names and restates one
reference function, cited as file:line under hhguo/MSMC-TTS `msmctts/`.  It is pinned by
oracle/make_golden.py: the unmodified reference, imported from /root/reference, is run on seeded inputs and
its outputs are committed under tests/golden/; tests/test_oracle_golden.py checks this file against them.
Parity status: PINNED for everything except MelLoss's mel filterbank, which the reference takes from librosa
(third-party, unpinned, absent) -- `slaney_mel_filterbank` restates librosa's published algorithm: PARITY UNPINNED.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def P(sd, prefix, name):
    return sd[prefix + name]


# ----------------------------------------------------------------------------------------------- utilities
def mask_from_lengths(lengths, max_len):
    """utils/utils.py:153-157 get_mask_from_lengths -- True where padded"""
    ids = torch.arange(0, max_len, device=lengths.device)
    return ~(ids < lengths.unsqueeze(1))


def weight_norm(sd, prefix, name="weight"):
    """torch.nn.utils.weight_norm, dim=0: w = g * v / ||v|| (norm over all dims but 0)"""
    if prefix + name in sd:
        return sd[prefix + name]
    v, g = sd[prefix + name + "_v"], sd[prefix + name + "_g"]
    dims = tuple(range(1, v.dim()))
    return v * (g / v.norm(2, dim=dims, keepdim=True))


def sinusoid_table(n_position, d_hid, padding_idx=0):
    """acoustic_models/transformer.py:388-408"""
    pos = np.arange(n_position)[:, None].astype(np.float64)
    j = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    if padding_idx is not None:
        tab[padding_idx] = 0.0
    return torch.FloatTensor(tab)


# ------------------------------------------------------------------------------------------------------ VQ
def quantize(sd, prefix, x, lengths=None, training=False, update=True, decay=0.99, eps=1e-5):
    """vqgantts/modules.py:24-67 Quantize.forward.  Mutates sd[prefix+'embed'|'cluster_size'|'embed_avg'] in
    place when training and update (lines 35-57)."""
    embed = sd[prefix + "embed"]
    dim, n_embed = embed.shape
    flatten = x.reshape(-1, dim)
    dist = flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ embed + embed.pow(2).sum(0, keepdim=True)
    _, ind = (-dist).max(1)
    ind = ind.view(*x.shape[:-1])
    q = F.embedding(ind, embed.transpose(0, 1))
    if training and update:
        with torch.no_grad():
            onehot = F.one_hot(ind, n_embed).type(flatten.dtype)
            onehot = torch.cat([onehot[i, : int(lengths[i])] for i in range(x.shape[0])], dim=0)
            flat = torch.cat([x[i, : int(lengths[i])] for i in range(x.shape[0])], dim=0)
            onehot_sum = onehot.sum(0)
            embed_sum = flat.transpose(0, 1) @ onehot
            cs, ea = sd[prefix + "cluster_size"], sd[prefix + "embed_avg"]
            cs.mul_(decay).add_(onehot_sum, alpha=1 - decay)
            ea.mul_(decay).add_(embed_sum, alpha=1 - decay)
            n = cs.sum()
            csn = (cs + eps) / (n + n_embed * eps) * n
            embed.copy_(ea / csn.unsqueeze(0))
    diff = (q.detach() - x).pow(2)
    q = x + (q - x).detach()
    return q, diff, ind


def multihead_quantize(sd, prefix, x, n_head, lengths=None, training=False, update=True):
    """vqgantts/modules.py:137-151 MultiHeadQuantize.forward"""
    heads = torch.chunk(x, n_head, dim=-1)
    qs, ds, inds = [], [], []
    for h, head in enumerate(heads):
        q, d, i = quantize(sd, "%squantizers.%d." % (prefix, h), head, lengths, training, update)
        qs.append(q); ds.append(d); inds.append(i)
    return torch.cat(qs, dim=-1), sum(ds) / len(ds), torch.stack(inds, dim=-1)


def any_quantize(sd, prefix, x, n_head, lengths, training, update):
    if n_head == 1:
        return quantize(sd, prefix, x, lengths, training, update)
    return multihead_quantize(sd, prefix, x, n_head, lengths, training, update)


def triple_loss(sd, prefix, pred, target, reduction="mean", margin=1e-6):
    """vqgantts/modules.py:86-116 Quantize.compute_triple_loss"""
    embed = sd[prefix + "embed"]
    dim = embed.shape[0]
    B, T, _ = pred.shape
    flatten = pred.reshape(-1, dim)
    dist = (flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ embed
            + embed.pow(2).sum(0, keepdim=True)).reshape(B, T, -1)
    pos = F.mse_loss(pred, F.embedding(target, embed.transpose(0, 1)), reduction="none").sum(-1)
    tl = pos.unsqueeze(-1) - dist
    mask = tl != 0
    tl = torch.clamp(tl + margin, min=0)
    tl = mask * (tl / dim)
    return tl.mean(-1) if reduction == "mean" else tl.sum(-1)


def multihead_triple_loss(sd, prefix, pred, target, n_head, reduction="mean"):
    """vqgantts/modules.py:153-169"""
    if n_head == 1:
        return triple_loss(sd, prefix, pred, target, reduction)
    ps = torch.chunk(pred, n_head, dim=-1)
    ts = torch.chunk(target, n_head, dim=-1)
    out = [triple_loss(sd, "%squantizers.%d." % (prefix, h), p, t.squeeze(-1), reduction)
           for h, (p, t) in enumerate(zip(ps, ts))]
    return sum(out) / len(out)


# ---------------------------------------------------------------------------------------------- FFT blocks
def multi_head_attention(sd, prefix, x, key_pad, n_head, d_k, d_v, p_drop=0.0, p_attn=0.0, training=False):
    """acoustic_models/transformer.py:246-328 MultiHeadAttention + ScaledDotProductAttention"""
    bs, t, _ = x.shape
    residual = x
    d_out = d_k + d_k + d_v
    y = F.linear(x, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    y = y.view(bs, t, n_head, d_out).permute(2, 0, 1, 3).contiguous().view(n_head * bs, t, d_out)
    q, k, v = y[..., :d_k], y[..., d_k:2 * d_k], y[..., 2 * d_k:]
    mask = key_pad.unsqueeze(1).expand(-1, t, -1).repeat(n_head, 1, 1)
    attn = torch.bmm(q, k.transpose(1, 2)) / np.power(d_k, 0.5)
    attn = attn.masked_fill(mask, -np.inf)
    attn = F.softmax(attn, dim=2)
    attn = F.dropout(attn, p_attn, training)
    out = torch.bmm(attn, v)
    out = out.view(n_head, bs, t, d_v).permute(1, 2, 0, 3).contiguous().view(bs, t, n_head * d_v)
    out = F.linear(out, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])
    out = F.dropout(out, p_drop, training) + residual
    return F.layer_norm(out, (out.shape[-1],), sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"])


def positionwise_ffn(sd, prefix, x, padding, p_drop=0.0, training=False):
    """acoustic_models/transformer.py:362-385 PositionwiseFeedForward"""
    residual = x
    y = F.conv1d(x.transpose(1, 2), sd[prefix + "w_1.weight"], sd[prefix + "w_1.bias"], padding=padding)
    y = F.relu(y)
    y = F.conv1d(y, sd[prefix + "w_2.weight"], sd[prefix + "w_2.bias"], padding=padding).transpose(1, 2)
    y = F.dropout(y, p_drop, training) + residual
    return F.layer_norm(y, (y.shape[-1],), sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"])


def fft_blocks(sd, prefix, seq, pos, cfg, training=False, use_dropout=False):
    """acoustic_models/transformer.py:119-146 FFTBlocks.forward + 195-206 FFTBlock.forward"""
    key_pad = pos.eq(0)
    non_pad = pos.ne(0).unsqueeze(-1).to(seq.dtype)
    table = sd[prefix + "position.weight"]
    out = seq + F.embedding(pos, table)
    pd = cfg["dropout"] if use_dropout else 0.0
    pa = cfg.get("attn_dropout", 0.1) if use_dropout else 0.0
    for i in range(cfg["n_layers"]):
        lp = "%slayer_stack.%d." % (prefix, i)
        out = multi_head_attention(sd, lp + "slf_attn.", out, key_pad, cfg["n_head"], cfg["d_k"], cfg["d_v"],
                                   pd, pa, training)
        out = out * non_pad
        out = positionwise_ffn(sd, lp + "pos_ffn.", out, cfg["fft_conv1d_padding"], pd, training)
        out = out * non_pad
    return out


def make_pos(lengths, max_len):
    """msmc_vqgan.py:56-58 -- positions 1..T, 0 on padding"""
    pos = torch.arange(1, max_len + 1, device=lengths.device).view(1, -1).repeat(lengths.shape[0], 1).long()
    pos.masked_fill_(mask_from_lengths(lengths, max_len), 0)
    return pos


# ------------------------------------------------------------------------------------- ResStack / predictor
def res_stack(sd, prefix, x, x_mask, hidden, kernel_size, dilation_rate, n_layers, p_drop=0.0, training=False):
    """vqgantts/modules.py:223-251 ResStack.forward (g is None on this path); x (B, C, T)"""
    output = torch.zeros_like(x)
    for i in range(n_layers):
        dilation = dilation_rate ** i
        padding = int((kernel_size * dilation - dilation) / 2)
        w = weight_norm(sd, "%sin_layers.%d." % (prefix, i))
        x_in = F.conv1d(x, w, sd["%sin_layers.%d.bias" % (prefix, i)], dilation=dilation, padding=padding)
        acts = torch.tanh(x_in[:, :hidden]) * torch.sigmoid(x_in[:, hidden:])
        acts = F.dropout(acts, p_drop, training)
        w = weight_norm(sd, "%sres_skip_layers.%d." % (prefix, i))
        rs = F.conv1d(acts, w, sd["%sres_skip_layers.%d.bias" % (prefix, i)])
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * x_mask
            output = output + rs[:, hidden:]
        else:
            output = output + rs
    return output * x_mask


def prior_predictor(sd, prefix, x, lengths, in_channels, prior_cfg, p_drop=0.0, training=False):
    """vqgantts/msmc_vqgan.py:82-88 PriorPredictor.forward"""
    x = x.transpose(1, 2)
    x_mask = (~mask_from_lengths(lengths, x.shape[-1])).unsqueeze(1).to(x.dtype)
    h = res_stack(sd, prefix + "enc.", x, x_mask, in_channels, prior_cfg.get("kernel_size", 5),
                  prior_cfg.get("dilation_rate", 1), prior_cfg.get("n_layers", 4), p_drop, training)
    o = F.conv1d(h, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"]) * x_mask
    return h.transpose(1, 2), o.transpose(1, 2)


# ------------------------------------------------------------------------------------------- MSMC-VQ-GAN
def multistage_encoder(sd, prefix, x, lengths, enc_cfg, training=False, use_dropout=False):
    """vqgantts/msmc_vqgan.py:46-62 MultiStageEncoder.forward"""
    outs = []
    feat, feat_len = x, lengths
    for i, scale in enumerate(enc_cfg["downsample_scales"]):
        if scale > 1:
            feat = F.avg_pool1d(feat.transpose(1, 2), kernel_size=scale, stride=scale, ceil_mode=True).transpose(1, 2)
            feat_len = torch.ceil(feat_len / scale).int()
        pos = make_pos(feat_len, feat.shape[1])
        feat = fft_blocks(sd, "%sencoders.%d." % (prefix, i), feat, pos, enc_cfg, training, use_dropout)
        outs.append((feat, feat_len))
    return outs


def multistage_quantizer(sd, prefix, encoder_states, n_model_size, upsample_scales, q_cfg, training=False,
                         use_dropout=False, from_encoder=True):
    """vqgantts/msmc_vqgan.py:147-234 MultiStageQuantizer.forward (upsampling == 'repeat')"""
    n_heads = q_cfg.get("n_heads", 4)
    p = q_cfg.get("dropout", 0.1) if use_dropout else 0.0
    prior_cfg = q_cfg.get("prior_config", {})
    update = q_cfg.get("update_codebook", True)
    quant_states, pred_states = [], []
    residual = None
    if from_encoder:
        encoder_states = encoder_states[::-1]
    for i, (emb, length) in enumerate(encoder_states):
        if residual is None:
            pred_quant = None
        else:
            residual = residual[:, : int(length.max())]
            pred_hidden, pred_quant = prior_predictor(sd, "%spredictor.%d." % (prefix, i), residual, length,
                                                      n_model_size, prior_cfg, p, training)
            residual = residual + F.dropout(pred_hidden, p, training)
        if emb is None:
            q_in = pred_quant
        elif from_encoder:
            pre = torch.cat((emb, residual), dim=-1) if residual is not None else emb
            pp = "%spreprocessor.%d." % (prefix, i)
            h = F.conv1d(pre.transpose(1, 2), sd[pp + "0.weight"], sd[pp + "0.bias"])
            h = torch.tanh(h)
            q_in = F.conv1d(h, sd[pp + "2.weight"], sd[pp + "2.bias"]).transpose(1, 2)
        else:
            q_in = emb
        quant, diffs, indices = any_quantize(sd, "%squantizer.%d." % (prefix, i), q_in, n_heads, length, training,
                                             update)
        post_in = quant if residual is None else torch.cat((residual, quant), dim=-1)
        po = "%spostprocessor.%d." % (prefix, i)
        h = torch.tanh(F.linear(post_in, sd[po + "0.weight"], sd[po + "0.bias"]))
        post_out = F.dropout(F.linear(h, sd[po + "2.weight"], sd[po + "2.bias"]), p, training)
        residual = post_out if residual is None else residual + post_out
        quant_states.append((quant, diffs, indices))
        pred_states.append(dict(predictor_outputs=pred_quant, target_outputs=quant, target_indices=indices,
                                target_lengths=length))
        residual = torch.repeat_interleave(residual, upsample_scales[i], dim=1)
    qo, qd, qi = zip(*quant_states)
    out = dict(residual_output=residual, quantizer_outputs=qo, quantizer_diffs=qd, quantizer_indices=qi,
               quantizer_lengths=[x[1] for x in encoder_states])
    out["predictor_diffs"] = embedding_loss(sd, prefix, pred_states, n_heads) if training else None
    return out


def embedding_loss(sd, prefix, pred_states, n_heads, methods=("mse",), loss_weights=(1.0,)):
    """vqgantts/msmc_vqgan.py:236-273 compute_embedding_loss"""
    loss_dict = {"total_loss": 0}
    for i, st in enumerate(pred_states):
        p = st["predictor_outputs"]
        if p is None:
            continue
        weights = loss_weights[i] if isinstance(loss_weights[0], (list, tuple)) else loss_weights
        for method, weight in zip(methods, weights):
            if method == "mse":
                loss = F.mse_loss(p, st["target_outputs"].detach(), reduction="none").mean(-1)
            elif method in ("triple", "triple_mean"):
                loss = multihead_triple_loss(sd, "%squantizer.%d." % (prefix, i), p, st["target_indices"], n_heads)
            elif method == "triple_sum":
                loss = multihead_triple_loss(sd, "%squantizer.%d." % (prefix, i), p, st["target_indices"], n_heads,
                                             "sum")
            else:
                raise ValueError(method)
            mask = mask_from_lengths(st["target_lengths"], loss.shape[1])
            loss = loss.masked_fill(mask, 0)
            loss = loss.sum() / sum(st["target_lengths"])
            loss_dict["embed_loss_%s_%d" % (method, i)] = loss
            loss_dict["total_loss"] = loss_dict["total_loss"] + loss * weight
    return loss_dict


def msmcvqgan_forward(sd, cfg, mel, mel_length, warmup=False, window=None, training=False, use_dropout=False,
                      prefix=""):
    """vqgantts/msmc_vqgan.py:309-350 MSMCVQGAN.forward.  cfg = the yaml `autoencoder` block."""
    n = cfg["n_model_size"]
    enc_cfg = cfg["encoder_config"]
    out = {}
    x = F.linear(mel, sd[prefix + "in_linear.weight"], sd[prefix + "in_linear.bias"])
    states = multistage_encoder(sd, prefix + "encoder.", x, mel_length, enc_cfg, training, use_dropout)
    qs = multistage_quantizer(sd, prefix + "quantizer.", states, n, enc_cfg["downsample_scales"][::-1],
                              cfg["quantizer_config"], training, use_dropout)
    dec_in = qs["residual_output"]
    eo, el = zip(*states)
    out.update(encoder_outputs=eo[::-1], encoder_lengths=el[::-1], encoder_indices=qs["quantizer_indices"],
               encoder_diffs=qs["quantizer_diffs"], decoder_diffs=qs["predictor_diffs"],
               quantizer_outputs=qs["quantizer_outputs"])
    if cfg.get("frame_decoder_config") is not None:
        pos = make_pos(mel_length, mel.shape[1])
        dec_in = fft_blocks(sd, prefix + "frame_decoder.", dec_in, pos, cfg["frame_decoder_config"], training,
                            use_dropout)
    if cfg.get("pred_mel", False):
        out["mel_outputs"] = F.linear(dec_in, sd[prefix + "mel_predictor.weight"], sd[prefix + "mel_predictor.bias"])
    if not warmup:
        if window is not None:
            dec_in = torch.stack([dec_in[i, s:e] for i, (s, e) in enumerate(window)], dim=0)
        out["decoder_outputs"] = generator(sd, prefix + "decoder.", dec_in.transpose(1, 2),
                                           cfg["decoder_config"]).transpose(1, 2)
    return out


# ------------------------------------------------------------------------------------ multi-stage predictor
def duration_predictor(sd, prefix, x, mask, p_drop=0.0, training=False):
    """acoustic_models/transformer.py:521-534 DurationPredictor.forward"""
    m = mask.to(x.dtype)
    out = x * m
    out = F.conv1d(out.transpose(1, 2), sd[prefix + "conv1d_1.weight"], sd[prefix + "conv1d_1.bias"], padding=1)
    out = F.relu(out.transpose(1, 2))
    out = F.layer_norm(out, (out.shape[-1],), sd[prefix + "layer_norm_1.weight"], sd[prefix + "layer_norm_1.bias"])
    out = F.dropout(out, p_drop, training)
    out = F.conv1d(out.transpose(1, 2), sd[prefix + "conv1d_2.weight"], sd[prefix + "conv1d_2.bias"], padding=1)
    out = F.relu(out.transpose(1, 2))
    out = F.layer_norm(out, (out.shape[-1],), sd[prefix + "layer_norm_2.weight"], sd[prefix + "layer_norm_2.bias"])
    out = F.dropout(out, p_drop, training)
    out = F.linear(out, sd[prefix + "linear_layer.weight"], sd[prefix + "linear_layer.bias"])
    return (out * m).squeeze(-1)


def multistage_predictor(sd, cfg, text, text_length, dur, feat, feat_length, training=True, prefix=""):
    """acoustic_models/multi_stage_predictor.py:43-126 MultiStagePredictor.forward (teacher-forced: dur and feat
    given, training mode, dropout off)"""
    n_symbols, scales = cfg["n_symbols"], cfg["n_pred_scale"]
    if isinstance(n_symbols, (list, tuple)):
        out = sum(F.embedding(text[..., i].long(), sd["%sword_emb.%d.weight" % (prefix, i)], padding_idx=0)
                  for i in range(len(n_symbols)))
    else:
        out = F.embedding(text.long(), sd[prefix + "word_emb.weight"], padding_idx=0)
    pos = make_pos(text_length, text.shape[1])
    out = fft_blocks(sd, prefix + "encoder.", out, pos, cfg["encoder_config"], training)
    text_mask = pos.ne(0).unsqueeze(-1)
    duration = duration_predictor(sd, prefix + "upsampler.duration_predictor.", out, text_mask)
    reps = torch.round(dur.float()).long()
    seqs = [torch.repeat_interleave(out[i], reps[i], dim=0) for i in range(out.shape[0])]
    emb = torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True)
    down = []
    for i, scale in enumerate(scales[::-1]):
        emb = F.conv1d(emb.transpose(1, 2), sd["%sdownsamplers.%d.weight" % (prefix, i)],
                       sd["%sdownsamplers.%d.bias" % (prefix, i)], padding=scale)
        emb = F.avg_pool1d(emb, kernel_size=scale, stride=scale, ceil_mode=True).transpose(1, 2)
        down.append(emb)
    down = down[::-1]
    preds, output = [], None
    for i in range(len(scales)):
        te = down[i]
        pos = make_pos(feat_length[i], te.shape[1])
        if i > 0:
            pre = torch.cat((output, feat[i - 1]), dim=2)
            pre = torch.repeat_interleave(pre, scales[i - 1], dim=1)[:, : te.shape[1]]
            output = torch.cat((te, pre), dim=2)
        else:
            output = te
        dp = "%sdecoders.%d." % (prefix, i)
        output = F.linear(output, sd[dp + "0.weight"], sd[dp + "0.bias"])
        output = fft_blocks(sd, dp + "1.", output, pos, cfg["decoder_config"], training)
        preds.append(F.linear(output, sd[dp + "2.weight"], sd[dp + "2.bias"]))
    return preds, duration


# ------------------------------------------------------------------------------------------------ HifiGAN
def resblock1(sd, prefix, x, kernel_size, dilations):
    """hifigan/common.py:43-51 ResBlock1.forward"""
    for j, d in enumerate(dilations):
        xt = F.leaky_relu(x, 0.1)
        xt = F.conv1d(xt, weight_norm(sd, "%sconvs1.%d." % (prefix, j)), sd["%sconvs1.%d.bias" % (prefix, j)],
                      dilation=d, padding=int((kernel_size * d - d) / 2))
        xt = F.leaky_relu(xt, 0.1)
        xt = F.conv1d(xt, weight_norm(sd, "%sconvs2.%d." % (prefix, j)), sd["%sconvs2.%d.bias" % (prefix, j)],
                      padding=int((kernel_size - 1) / 2))
        x = xt + x
    return x


def generator(sd, prefix, mel, dcfg):
    """hifigan/generator.py:40-55 Generator.forward; mel (B, C, L) -> (B, 1, L*prod(upsample_rates))"""
    ks, ds = dcfg["resblock_kernel_sizes"], dcfg["resblock_dilation_sizes"]
    nk = len(ks)
    x = F.conv1d(mel, weight_norm(sd, prefix + "conv_pre."), sd[prefix + "conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(dcfg["upsample_rates"], dcfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, weight_norm(sd, "%sups.%d." % (prefix, i)), sd["%sups.%d.bias" % (prefix, i)],
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = resblock1(sd, "%sresblocks.%d." % (prefix, i * nk + j), x, ks[j], ds[j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (generator.py:52)
    x = F.conv1d(x, weight_norm(sd, prefix + "conv_post."), sd[prefix + "conv_post.bias"], padding=3)
    return torch.tanh(x)


# ------------------------------------------------------------------------------------------ discriminator
def create_fb_matrix(n_freqs, f_min, f_max, n_mels, sample_rate):
    """utils/audio.py:30-84 (norm=None): HTK triangles clamped to [1e-6, 1]"""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), 1e-6, 1)


def stft_real(x, n_fft, hop, win, window, center=True, normalized=False):
    """torch.stft with the legacy real-view semantics the reference relies on (audio.py:399-402): (B, F, frames, 2)"""
    return torch.view_as_real(torch.stft(x, n_fft, hop, win, window, center=center, pad_mode="reflect",
                                         normalized=normalized, onesided=True, return_complex=True))


def torch_stft_transform(x, hop, sample_rate=24000, domain="double", mel_scale=True, ref_db=20, min_db=-100):
    """utils/audio.py:398-419 TorchSTFT.transform with fft=win=4*hop, normalized=True (discriminator.py:86-90)
    followed by MelScale (audio.py:348-376) with n_mels == n_freqs.  x (B, L) -> (B, 2F | F, frames)"""
    n_fft = hop * 4
    st = stft_real(x, n_fft, hop, n_fft, torch.hann_window(n_fft, device=x.device), normalized=True)
    mag = torch.sqrt(torch.clamp(st[..., 0] ** 2 + st[..., 1] ** 2, min=1e-7))
    if mel_scale:
        nf = n_fft // 2 + 1
        fb = create_fb_matrix(nf, 0.0, float(sample_rate // 2), nf, sample_rate).to(mag.device)
        mag = torch.matmul(mag.transpose(1, 2), fb).transpose(1, 2)
    if domain == "linear":
        return mag
    log_mag = 20 * torch.log10(mag) - ref_db
    log_mag = torch.clamp((log_mag - min_db) / -min_db, 0, 1)
    if domain == "log":
        return log_mag
    return torch.cat((mag, log_mag), dim=1)


def discriminator_r(sd, prefix, x):
    """hifigan/discriminator.py:15-76 DiscriminatorR.forward.  NOTE the in-place LeakyReLU(0.2, True) at the head
    of layers 1..6 mutates the tensor already appended to `hiddens`, so the returned feature maps are
    POST-activation."""
    strides = [1, 2, 1, 2, 1, 2, 1]
    hiddens = []
    for i, s in enumerate(strides):
        if i > 0:
            x = F.leaky_relu(x, 0.2)
            hiddens[-1] = x
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        # Sequential index of the conv: layer 0 = [pad, conv] -> '1', layers 1.. = [lrelu, pad, conv] -> '2'
        cp = "%sdiscriminator.%d.%d." % (prefix, i, 1 if i == 0 else 2)
        x = F.conv2d(x, weight_norm(sd, cp), sd[cp + "bias"], stride=s)
        hiddens.append(x)
    return x, hiddens[:-1]


def mrd(sd, prefix, y, mrd_cfg):
    """hifigan/discriminator.py:101-116 MultiResolutionDiscriminator.forward; y (B, 1, L)"""
    scores, feats = [], []
    for i, hop in enumerate(mrd_cfg["hop_lengths"]):
        mag = torch_stft_transform(y.squeeze(1), hop, mrd_cfg.get("sample_rate", 24000), mrd_cfg.get("domain", "double"),
                                   mrd_cfg.get("mel_scale", True))
        if mrd_cfg.get("domain", "double") == "double":
            mag = torch.stack(torch.chunk(mag, 2, dim=1), dim=1)
        else:
            mag = mag.unsqueeze(1)
        s, f = discriminator_r(sd, "%sdiscriminators.%d." % (prefix, i), mag)
        scores.append(s); feats.append(f)
    return scores, feats


def discriminator_p(sd, prefix, x, period):
    """hifigan/discriminator.py:135-154 DiscriminatorP.forward"""
    fmap = []
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    for i in range(5):
        cp = "%sconvs.%d." % (prefix, i)
        stride = (3, 1) if i < 4 else (1, 1)
        x = F.conv2d(x, weight_norm(sd, cp), sd[cp + "bias"], stride=stride, padding=(2, 0))
        fmap.append(x)
        x = F.leaky_relu(x, 0.2)
    x = F.conv2d(x, weight_norm(sd, prefix + "conv_post."), sd[prefix + "conv_post.bias"], padding=(1, 0))
    return torch.flatten(x, 1, -1), fmap


def discriminator(sd, prefix, y, dcfg):
    """hifigan/discriminator.py:180-190 Discriminator.forward"""
    if y.dim() == 2:
        y = y.unsqueeze(1)
    so, fo = mrd(sd, prefix + "mrd.", y, dcfg["mrd_config"])
    sp, fp = [], []
    for i, p in enumerate(dcfg["mpd_config"].get("periods", [2, 3, 5, 7, 11])):
        s, f = discriminator_p(sd, "%smpd.discriminators.%d." % (prefix, i), y, p)
        sp.append(s); fp.append(f)
    return so + sp, fo + fp


# ------------------------------------------------------------------------------------------------ MelLoss
def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(htk=False, norm='slaney') restated from librosa's published algorithm.
    PARITY UNPINNED: librosa is a third-party, unpinned dependency (requirements.txt:6) absent from the tree."""
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_pts = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels]))[:, None]
    return torch.from_numpy(weights.astype(np.float32))


def mel_spectrogram(y, fft_size, hop_size, win_size, sample_rate, num_mels):
    """trainers/criterions/stft_loss.py:78-107 MelLoss.mel_spectrogram"""
    basis = slaney_mel_filterbank(sample_rate, fft_size, num_mels, 0, sample_rate // 2).to(y.device)
    pad = int((fft_size - hop_size) / 2)
    y = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = stft_real(y, fft_size, hop_size, win_size, torch.hann_window(win_size, device=y.device), center=False)
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    spec = torch.matmul(basis, spec)
    return torch.log(torch.clamp(spec, min=1e-5))


def mel_loss(pred, target, sample_rate=24000, num_mels=128):
    """stft_loss.py:72-76 + msmctts_trainer.py:102-110 default kwargs"""
    win = sample_rate // 20
    hop = sample_rate // 80
    fft = 2048 if win > 1024 else 1024
    return F.l1_loss(mel_spectrogram(pred, fft, hop, win, sample_rate, num_mels),
                     mel_spectrogram(target, fft, hop, win, sample_rate, num_mels))


# --------------------------------------------------------------------------------------------- train step
def quantizer_loss(outputs, lambda_vq=1, lambda_pr=1):
    """trainers/msmctts_trainer.py:45-71 QuantizerLoss.forward (returns vq_loss and the named terms)"""
    loss = {"vq_loss": 0}
    for i, term in enumerate(outputs["encoder_diffs"]):
        length = outputs["encoder_lengths"][i]
        mask = mask_from_lengths(length, term.shape[1])
        term = term.masked_fill(mask.unsqueeze(-1), 0)
        term = term.sum() / sum(length) / term.shape[2]
        loss["latent_loss_%d_0" % i] = term
        loss["vq_loss"] = loss["vq_loss"] + lambda_vq * term
    dd = outputs.get("decoder_diffs")
    if isinstance(dd, dict):
        dd = dict(dd)
        loss["vq_loss"] = loss["vq_loss"] + lambda_pr * dd.pop("total_loss")
        loss.update(dd)
    return loss


def generator_losses(sd_ae, sd_d, cfg, mel, mel_length, wav, windows, tcfg, training=True, use_dropout=False):
    """trainers/msmctts_trainer.py:115-201: everything the G step differentiates (vq + frame + stft + adv + fm)
    and the D loss on the same forward.  windows = list[(start_frame, end_frame)]."""
    fs = tcfg.get("frameshift", 300)
    target = torch.stack([wav[i, s * fs:e * fs] for i, (s, e) in enumerate(windows)], dim=0).squeeze(-1)
    out = msmcvqgan_forward(sd_ae, cfg["autoencoder"], mel, mel_length, False, windows, training, use_dropout)
    vq = quantizer_loss(out, tcfg.get("lambda_vq", 1), tcfg.get("lambda_pr", 1))
    g_loss = vq["vq_loss"]
    losses = {"vq_loss": vq["vq_loss"]}
    if "mel_outputs" in out:
        ml = F.mse_loss(mel, out["mel_outputs"], reduction="none")
        ml = ml.masked_fill(mask_from_lengths(mel_length, mel.shape[1]).unsqueeze(-1), 0)
        ml = ml.sum() / sum(mel_length) / ml.shape[2]
        losses["frame_loss"] = ml
        g_loss = g_loss + tcfg.get("lambda_frame", 1.0) * ml
    predict = out["decoder_outputs"].squeeze(-1)
    stft = mel_loss(predict, target, tcfg.get("sample_rate", 24000))
    losses["stft_loss"] = stft
    g_loss = g_loss + tcfg.get("lambda_stft", 45) * stft
    # discriminator step (lines 162-175)
    fs_d, _ = discriminator(sd_d, "", predict.detach(), cfg["discriminator"])
    rs_d, _ = discriminator(sd_d, "", target, cfg["discriminator"])
    d_real = sum(F.mse_loss(r, torch.ones_like(r)) for r in rs_d)
    d_fake = sum(F.mse_loss(f, torch.zeros_like(f)) for f in fs_d)
    losses.update(d_loss_real=d_real, d_loss_fake=d_fake, d_loss=d_real + d_fake)
    # generator step (lines 182-201); NOTE in the reference the D weights have already been stepped here
    return losses, g_loss, predict, target, out


def generator_adv_losses(sd_d, cfg, predict, target, g_loss, lambda_fm=2):
    """trainers/msmctts_trainer.py:182-201"""
    fs_g, ff = discriminator(sd_d, "", predict, cfg["discriminator"])
    _, rf = discriminator(sd_d, "", target, cfg["discriminator"])
    adv = sum(F.mse_loss(f, torch.ones_like(f)) for f in fs_g)
    fm = sum(F.l1_loss(a, b) for fa, fb in zip(ff, rf) for a, b in zip(fa, fb))
    adv_total = adv + fm * lambda_fm
    return dict(fm_loss=fm, adv_loss=adv_total, g_loss=g_loss + adv_total)
