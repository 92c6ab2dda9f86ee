"""VQ quantisers and the WaveNet-style ResStack on the sm_100a kernels.  Class names, kwargs, buffers and
state_dict keys follow the reference's vqgantts/modules.py (Quantize :10-116, MultiHeadQuantize :119-169,
ResStack :182-260); forward returns the same (quantize, diff, embed_ind) triple.

All heads of a stage are searched by ONE launch of msmc_vq_search (the reference loops over heads in Python,
modules.py:141-146) and the EMA update is two launches (masked count/sum + renormalise) instead of the
reference's per-sample Python slicing with 2*B host syncs per head (modules.py:38-41).
"""
import torch
from torch import nn

from msmctts._b200 import functional as Fn
from msmctts._b200 import layers as Ly
from msmctts.utils.utils import get_mask_from_lengths, lengths_i32


class Quantize(nn.Module):
    def __init__(self, embed_dim, n_embed, decay=0.99, eps=1e-5):
        super().__init__()
        self.dim, self.n_embed, self.decay, self.eps = embed_dim, n_embed, decay, eps
        embed = torch.randn(embed_dim, n_embed)
        self.register_buffer("embed", embed)
        self.register_buffer("cluster_size", torch.zeros(n_embed))
        self.register_buffer("embed_avg", embed.clone())

    def _stacked(self):
        return self.embed.unsqueeze(0), self.embed_avg.unsqueeze(0), self.cluster_size.unsqueeze(0)

    def forward(self, input, input_length=None, update=True, sort=False):
        if sort:
            raise NotImplementedError("sort=True (full distance ranking) is not on the training hot path")
        embed, embed_avg, cluster_size = self._stacked()
        quant, diff, ind = Fn.vq_quantize(input, embed, 1, self.dim)
        if self.training and update:
            Fn.vq_ema_update(input, ind, lengths_i32(input_length, input.device), embed, embed_avg, cluster_size,
                             self.decay, self.eps)
        return quant, diff, ind.squeeze(-1)

    def embed_code(self, embed_id):
        return torch.nn.functional.embedding(embed_id, self.embed.transpose(0, 1))

    def compute_triple_loss(self, prd_quant, trg_quant, reduction="mean", margin=1e-6, adaptive_margin=False):
        return Fn.vq_triple_loss(prd_quant, self.embed.unsqueeze(0), trg_quant.unsqueeze(-1), 1, self.dim, margin,
                                 reduction)


class MultiHeadQuantize(nn.Module):
    def __init__(self, embed_dim, n_embed, n_head, decay=0.99, eps=1e-5):
        super().__init__()
        assert embed_dim % n_head == 0
        self.dim, self.n_embed, self.n_head, self.decay, self.eps = embed_dim, n_embed, n_head, decay, eps
        self.sub_dim = embed_dim // n_head
        self.quantizers = nn.ModuleList([Quantize(self.sub_dim, n_embed, decay, eps) for _ in range(n_head)])
        self._stack = None

    def _stacked(self):
        """Per-head buffers are views into one (n_head, dim, K) block so a single launch serves every head while
        state_dict() still exposes quantizers.{h}.embed / cluster_size / embed_avg.  Re-established whenever
        .to()/.cuda()/load_state_dict replaced the buffers."""
        q0 = self.quantizers[0]
        st = self._stack
        ok = st is not None and st[0].device == q0.embed.device and all(
            q.embed.data_ptr() == st[0][h].data_ptr() and q.embed_avg.data_ptr() == st[1][h].data_ptr()
            and q.cluster_size.data_ptr() == st[2][h].data_ptr() for h, q in enumerate(self.quantizers))
        if not ok:
            with torch.no_grad():
                e = torch.stack([q.embed for q in self.quantizers]).contiguous()
                a = torch.stack([q.embed_avg for q in self.quantizers]).contiguous()
                c = torch.stack([q.cluster_size for q in self.quantizers]).contiguous()
            for h, q in enumerate(self.quantizers):
                q._buffers["embed"], q._buffers["embed_avg"], q._buffers["cluster_size"] = e[h], a[h], c[h]
            self._stack = st = (e, a, c)
        return st

    def forward(self, input, input_length=None, update=True, sort=False):
        if sort:
            raise NotImplementedError("sort=True is not on the training hot path")
        embed, embed_avg, cluster_size = self._stacked()
        quant, diff, ind = Fn.vq_quantize(input, embed, self.n_head, self.sub_dim)
        if self.training and update:
            Fn.vq_ema_update(input, ind, lengths_i32(input_length, input.device), embed, embed_avg, cluster_size,
                             self.decay, self.eps)
        return quant, diff, ind

    def compute_triple_loss(self, prd_quant, trg_quant, reduction="mean", margin=1e-6, adaptive_margin=False):
        embed, _, _ = self._stacked()
        return Fn.vq_triple_loss(prd_quant, embed, trg_quant, self.n_head, self.sub_dim, margin, reduction)


class ResStack(nn.Module):
    """WaveNet-style gated stack; x is (B, T, C) here, the reference's (B, C, T) (modules.py:223-251)."""

    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0.1):
        super().__init__()
        assert kernel_size % 2 == 1
        if gin_channels != 0:
            raise NotImplementedError("global conditioning is unused by the in-tree configs")
        self.hidden_channels, self.n_layers, self.p_dropout = hidden_channels, n_layers, p_dropout
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        self.drop = nn.Dropout(p_dropout)
        for i in range(n_layers):
            dilation = dilation_rate ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(Ly.WNConv1d(hidden_channels, 2 * hidden_channels, kernel_size, dilation=dilation,
                                              padding=padding))
            out_ch = 2 * hidden_channels if i < n_layers - 1 else hidden_channels
            self.res_skip_layers.append(Ly.WNConv1d(hidden_channels, out_ch, 1))

    def forward(self, x, x_mask, g=None, **kwargs):
        H = self.hidden_channels
        output = None
        for i in range(self.n_layers):
            acts = self.drop(Fn.gated_act(self.in_layers[i](x)))
            rs = self.res_skip_layers[i](acts)
            if i < self.n_layers - 1:
                x = (x + rs[..., :H]) * x_mask
                skip = rs[..., H:]
            else:
                skip = rs
            output = skip if output is None else output + skip
        return output * x_mask


class Encoder(nn.Module):
    """modules.py:263-289 (unused by the CSMSC configs; kept for API parity)"""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size=5, dilation_rate=1, n_layers=16):
        super().__init__()
        self.pre = Ly.Conv1d(in_channels, hidden_channels, 1)
        self.enc = ResStack(hidden_channels, kernel_size, dilation_rate, n_layers)
        self.proj = Ly.Conv1d(hidden_channels, out_channels, 1)

    def forward(self, x, x_lengths):
        x_mask = (~get_mask_from_lengths(x_lengths, x.shape[1])).unsqueeze(-1).to(x.dtype)
        h = self.enc(self.pre(x) * x_mask, x_mask)
        return self.proj(h) * x_mask, h
