"""FFT blocks on the sm_100a kernels.  Same classes, kwargs and state_dict keys as the reference's
acoustic_models/transformer.py (FFTBlocks :71-146, FFTBlock :149-206, MultiHeadAttention :209-286,
PositionwiseFeedForward :331-385, LengthRegulator :427-484, DurationPredictor :487-534); the arithmetic runs in
  QKV / fc Linear        -> msmc_conv_forward (native weight layout)
  masked softmax(QK^T)V  -> msmc_attention_fwd/bwd (no (t,t) matrix, no permute copies)
  dropout+residual+LN+mask -> msmc_add_layernorm_fwd/bwd
  conv-FFN k=3           -> msmc_conv_forward with ReLU fused in the epilogue
Sequences stay (B, t, C) (channels-last) throughout: the reference's NCL transposes disappear.
"""
import numpy as np
import torch
from torch import nn
from torch.nn.utils.rnn import pad_sequence

from msmctts._b200 import functional as Fn
from msmctts._b200 import layers as Ly


def get_sinusoid_encoding_table(n_position, d_hid, padding_idx=None):
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    if padding_idx is not None:
        table[padding_idx] = 0.0
    return torch.FloatTensor(table)


def get_non_pad_mask(seq):
    return seq.ne(0).unsqueeze(-1)


def get_attn_key_pad_mask(seq_k, seq_q):
    return seq_k.eq(0).unsqueeze(1).expand(-1, seq_q.size(1), -1)


def lengths_from_pos(pos):
    """positions are 1..len then zeros (padding is trailing by construction, msmc_vqgan.py:56-58)"""
    return pos.ne(0).sum(dim=1).to(torch.int32)


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature, attn_dropout=0.1, name=None):
        super().__init__()
        self.temperature = float(temperature)
        self.attn_dropout = attn_dropout
        self.name = name


class MultiHeadAttention(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout, name, attn_dropout=0.1, fused_layernorm=False):
        super().__init__()
        if d_k != d_v:
            raise ValueError("the fused attention kernel needs d_k == d_v")
        self.n_head, self.d_k, self.d_v, self.name = n_head, d_k, d_v, name
        self.linear = Ly.Linear(d_model, n_head * (2 * d_k + d_v))
        nn.init.xavier_normal_(self.linear.weight)
        self.attention = ScaledDotProductAttention(np.power(d_k, 0.5), attn_dropout, "%s.scaled_dot" % name)
        self.layer_norm = Ly.LayerNormParams(d_model)
        self.fc = Ly.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.p_dropout = dropout

    def forward(self, x, lengths):
        p_attn = self.attention.attn_dropout if self.training else 0.0
        p_out = self.p_dropout if self.training else 0.0
        qkv = self.linear(x)
        ctx = Fn.attention(qkv, lengths, self.n_head, self.d_k, self.attention.temperature, p_attn)
        out = self.fc(ctx)
        return self.layer_norm(out, x, lengths, p_out)          # mask(LN(dropout(out) + x))


class PositionwiseFeedForward(nn.Module):
    def __init__(self, d_in, d_hid, fft_conv1d_kernel, fft_conv1d_padding, dropout, name, fused_layernorm=False):
        super().__init__()
        self.name = name
        self.w_1 = Ly.Conv1d(d_in, d_hid, fft_conv1d_kernel, padding=fft_conv1d_padding)
        self.w_2 = Ly.Conv1d(d_hid, d_in, fft_conv1d_kernel, padding=fft_conv1d_padding)
        self.layer_norm = Ly.LayerNormParams(d_in)
        self.p_dropout = dropout

    def forward(self, x, lengths):
        h = self.w_1(x, post="relu")
        o = self.w_2(h)
        return self.layer_norm(o, x, lengths, self.p_dropout if self.training else 0.0)


class FFTBlock(nn.Module):
    def __init__(self, d_model, d_inner, n_head, d_k, d_v, fft_conv1d_kernel, fft_conv1d_padding, dropout, name,
                 attn_dropout=0.1, fused_layernorm=False):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout, "%s.slf_attn" % name, attn_dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, fft_conv1d_kernel, fft_conv1d_padding, dropout,
                                               "%s.pos_ffn" % name)

    def forward(self, x, lengths):
        return self.pos_ffn(self.slf_attn(x, lengths), lengths)


class FFTBlocks(nn.Module):
    def __init__(self, max_seq_len, n_layers, n_head, d_k, d_v, d_model, d_inner, fft_conv1d_kernel,
                 fft_conv1d_padding, dropout, name, attn_dropout=0.1, fused_layernorm=False):
        super().__init__()
        self.max_seq_len, self.n_layers, self.d_model, self.name = max_seq_len, n_layers, d_model, name
        self.position = nn.Embedding.from_pretrained(
            get_sinusoid_encoding_table(max_seq_len + 1, d_model, padding_idx=0), freeze=True)
        self.layer_stack = nn.ModuleList([
            FFTBlock(d_model, d_inner, n_head, d_k, d_v, fft_conv1d_kernel, fft_conv1d_padding, dropout,
                     "%s.layer_stack.%d" % (name, i), attn_dropout) for i in range(n_layers)])

    def forward(self, seq, pos, return_attns=False, acts=None):
        lengths = lengths_from_pos(pos)
        out = seq + self.position(pos)
        for layer in self.layer_stack:
            out = layer(out, lengths)
        return out, get_non_pad_mask(pos)


class DurationPredictor(nn.Module):
    """transformer.py:487-534 (conv k=3 pad=1 -> ReLU -> LN -> dropout) x2 -> Linear(1)"""

    def __init__(self, input_size, filter_size, kernel, dropout, fused_layernorm=False):
        super().__init__()
        self.dropout = dropout
        self.conv1d_1 = Ly.Conv1d(input_size, filter_size, kernel, padding=1)
        self.layer_norm_1 = Ly.LayerNormParams(filter_size)
        self.conv1d_2 = Ly.Conv1d(filter_size, filter_size, kernel, padding=1)
        self.layer_norm_2 = Ly.LayerNormParams(filter_size)
        self.linear_layer = Ly.Linear(filter_size, 1)

    def forward(self, x, mask):
        m = mask.to(x.dtype)
        p = self.dropout if self.training else 0.0
        out = self.conv1d_1(x * m, post="relu")
        out = torch.nn.functional.dropout(self.layer_norm_1(out), p, self.training)
        out = self.conv1d_2(out, post="relu")
        out = torch.nn.functional.dropout(self.layer_norm_2(out), p, self.training)
        return (self.linear_layer(out) * m).squeeze(-1)


class LengthRegulator(nn.Module):
    def __init__(self, input_size, duration_predictor_filter_size, duration_predictor_kernel_size, dropout,
                 fused_layernorm=False):
        super().__init__()
        self.duration_predictor = DurationPredictor(input_size, duration_predictor_filter_size,
                                                    duration_predictor_kernel_size, dropout)

    def forward(self, x, mask, target=None, alpha=1.0):
        duration = self.duration_predictor(x, mask)
        if self.training:
            out, pos = self.get_output(x, target, alpha)
            return out, pos, duration
        duration = torch.clamp_min(duration, 0) if target is None else target
        out, pos = self.get_output(x, duration, alpha)
        return out, pos, torch.round(duration).long()

    def get_output(self, x, duration, alpha):
        """x (B, L, C) expanded by per-token repeat counts -> (B, T, C), zero padded; positions 1..len (0 = pad).
        The reference loops over the batch (repeat_interleave + pad_sequence per sample, transformer.py:470-489);
        here one searchsorted + gather serves the whole batch.  The output length is data dependent, so ONE host
        read of max(total) remains (the reference has B of them)."""
        B, Ln, Cn = x.shape
        reps = torch.round(duration.float() * alpha).long().clamp_min(0)
        csum = reps.cumsum(1)
        total = csum[:, -1]
        T = int(total.max())
        t = torch.arange(T, device=x.device).unsqueeze(0).expand(B, T)
        idx = torch.searchsorted(csum, t.contiguous(), right=True).clamp_max(Ln - 1)
        valid = t < total.unsqueeze(1)
        out = torch.gather(x, 1, idx.unsqueeze(-1).expand(-1, -1, Cn)) * valid.unsqueeze(-1).to(x.dtype)
        pos = torch.where(valid, t + 1, torch.zeros_like(t))
        return out, pos
