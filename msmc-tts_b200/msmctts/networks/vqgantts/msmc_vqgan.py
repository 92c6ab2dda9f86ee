"""MSMC-VQ-GAN autoencoder on the sm_100a kernels.  Same classes / kwargs / forward signatures / state_dict keys as
the reference's vqgantts/msmc_vqgan.py (MultiStageEncoder :14-62, PriorPredictor :65-88, MultiStageQuantizer
:91-273, MSMCVQGAN :276-410).  Everything stays (B, T, C) channels-last; no `.item()`/`int(tensor)` host syncs on
the forward path (the reference has ~275 per step), so a whole train step is CUDA-graph capturable.
"""
import torch
from torch import nn
from torch.nn import functional as F

from msmctts._b200 import layers as Ly
from msmctts.networks.acoustic_models.transformer import FFTBlocks
from msmctts.networks.hifigan import HifiGANGenerator
from msmctts.utils.utils import get_mask_from_lengths
from .modules import MultiHeadQuantize, Quantize, ResStack


def _positions(lengths, t):
    """1..len then 0 on padding (reference builds this with a Python list + range(length.max()), :56-58)"""
    ids = torch.arange(1, t + 1, device=lengths.device).unsqueeze(0)
    return torch.where(ids <= lengths.unsqueeze(1), ids, torch.zeros_like(ids)).long()


class MultiStageEncoder(nn.Module):
    def __init__(self, in_channels, downsample_scales=[1], max_seq_len=2400, n_layers=4, n_head=2, d_k=64, d_v=64,
                 d_inner=1024, fft_conv1d_kernel=3, fft_conv1d_padding=1, dropout=0.2, attn_dropout=0.1,
                 fused_layernorm=False):
        super().__init__()
        self.downsample_scales = list(downsample_scales)
        self.encoders = nn.ModuleList([
            FFTBlocks(max_seq_len=max_seq_len, n_layers=n_layers, n_head=n_head, d_k=d_k, d_v=d_v,
                      d_model=in_channels, d_inner=d_inner, fft_conv1d_kernel=fft_conv1d_kernel,
                      fft_conv1d_padding=fft_conv1d_padding, dropout=dropout, attn_dropout=attn_dropout,
                      name="encoder_%d" % i) for i in range(len(self.downsample_scales))])

    def forward(self, input, input_length):
        outputs = []
        feat, feat_length = input, input_length
        for encoder, scale in zip(self.encoders, self.downsample_scales):
            if scale > 1:
                feat = F.avg_pool1d(feat.transpose(1, 2), kernel_size=scale, stride=scale,
                                    ceil_mode=True).transpose(1, 2)
                feat_length = torch.ceil(feat_length / scale).int()
            feat, _ = encoder(feat, _positions(feat_length, feat.shape[1]))
            outputs.append((feat, feat_length))
        return outputs


class PriorPredictor(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=5, dilation_rate=1, n_layers=4):
        super().__init__()
        self.enc = ResStack(in_channels, kernel_size, dilation_rate, n_layers)
        self.proj = Ly.Conv1d(in_channels, out_channels, 1)

    def forward(self, x, x_lengths):
        x_mask = (~get_mask_from_lengths(x_lengths, x.shape[1])).unsqueeze(-1).to(x.dtype)
        h = self.enc(x, x_mask)
        return h, self.proj(h) * x_mask


class _Seq(nn.Module):
    """children named '0' and '2' like the reference's nn.Sequential(conv|linear, Tanh, conv|linear)"""

    def __init__(self, first, second):
        super().__init__()
        self.add_module("0", first)
        self.add_module("2", second)

    def forward(self, x):
        return getattr(self, "2")(getattr(self, "0")(x, post="tanh"))


class MultiStageQuantizer(nn.Module):
    def __init__(self, n_model_size, upsample_scales, embedding_sizes=512, embedding_dims=256, n_heads=4,
                 prior_config={}, norm=False, upsampling="repeat", dropout=0.1, update_codebook=True):
        super().__init__()
        if upsampling != "repeat" or norm:
            raise NotImplementedError("only upsampling='repeat', norm=False (the in-tree configs) are built")
        self.upsample_scales, self.upsampling = list(upsample_scales), upsampling
        self.dropout, self.update_codebook = dropout, update_codebook
        self.quantizer = nn.ModuleList()
        self.predictor = nn.ModuleList()
        self.preprocessor = nn.ModuleList()
        self.postprocessor = nn.ModuleList()
        for i in range(len(self.upsample_scales)):
            self.predictor.append(PriorPredictor(n_model_size, embedding_dims, **prior_config))
            self.preprocessor.append(_Seq(Ly.Conv1d(n_model_size * (1 if i == 0 else 2), embedding_dims, 1),
                                          Ly.Conv1d(embedding_dims, embedding_dims, 1)))
            self.quantizer.append(Quantize(embedding_dims, embedding_sizes) if n_heads == 1 else
                                  MultiHeadQuantize(embedding_dims, embedding_sizes, n_heads))
            self.postprocessor.append(_Seq(Ly.Linear(embedding_dims * (1 if i == 0 else 2), embedding_dims),
                                           Ly.Linear(embedding_dims, n_model_size)))

    def forward(self, encoder_states, from_encoder=True):
        quant_states, pred_states = [], []
        residual = None
        encoder_states = list(encoder_states)
        if from_encoder:
            encoder_states = encoder_states[::-1]
        for i, (embedding, length) in enumerate(encoder_states):
            if residual is None:
                pred_quant = None
            else:
                pred_hidden, pred_quant = self.predictor[i](residual, length)
                residual = residual + F.dropout(pred_hidden, p=self.dropout, training=self.training)
            if embedding is None:
                q_in = pred_quant
            elif from_encoder:
                pre = torch.cat((embedding, residual), dim=-1) if residual is not None else embedding
                q_in = self.preprocessor[i](pre)
            else:
                q_in = embedding
            quant, diffs, indices = self.quantizer[i](q_in, length, update=self.update_codebook)
            post_in = quant if residual is None else torch.cat((residual, quant), dim=-1)
            post_out = F.dropout(self.postprocessor[i](post_in), p=self.dropout, training=self.training)
            residual = post_out if residual is None else residual + post_out
            quant_states.append((quant, diffs, indices))
            pred_states.append({"predictor_outputs": pred_quant, "target_outputs": quant,
                                "target_indices": indices, "target_lengths": length})
            residual = torch.repeat_interleave(residual, self.upsample_scales[i], dim=1)
            if i + 1 < len(encoder_states) and encoder_states[i + 1][0] is not None:
                residual = residual[:, : encoder_states[i + 1][0].shape[1]]
        qo, qd, qi = zip(*quant_states)
        out = {"residual_output": residual, "quantizer_outputs": qo, "quantizer_diffs": qd,
               "quantizer_indices": qi, "quantizer_lengths": [x[1] for x in encoder_states]}
        out["predictor_diffs"] = self.compute_embedding_loss(pred_states, ["mse"], [1.0]) if self.training else None
        return out

    def compute_embedding_loss(self, pred_states, methods=["mse"], loss_weights=[1.0]):
        loss_dict = {"total_loss": 0}
        for i, state in enumerate(pred_states):
            p = state["predictor_outputs"]
            if p is None:
                continue
            weights = loss_weights[i] if isinstance(loss_weights[0], (list, tuple)) else loss_weights
            for method, weight in zip(methods, weights):
                if method == "mse":
                    loss = F.mse_loss(p, state["target_outputs"].detach(), reduction="none").mean(-1)
                elif method in ("triple", "triple_mean"):
                    loss = self.quantizer[i].compute_triple_loss(p, state["target_indices"])
                elif method == "triple_sum":
                    loss = self.quantizer[i].compute_triple_loss(p, state["target_indices"], reduction="sum")
                elif method == "softmax":
                    B, T, D = p.shape
                    loss = F.cross_entropy(p.reshape(-1, D), state["target_indices"].detach().reshape(-1),
                                           reduction="none").view(B, T)
                else:
                    raise ValueError(method)
                lengths = state["target_lengths"]
                loss = loss.masked_fill(get_mask_from_lengths(lengths, loss.shape[1]), 0)
                loss = loss.sum() / lengths.sum()
                loss_dict["embed_loss_{}_{}".format(method, i)] = loss
                loss_dict["total_loss"] = loss_dict["total_loss"] + loss * weight
        return loss_dict


class MSMCVQGAN(nn.Module):
    def __init__(self, in_dim, n_model_size, encoder_config=None, quantizer_config=None,
                 frame_decoder_config=None, decoder_config=None, pred_mel=False):
        super().__init__()
        self.in_linear = Ly.Linear(in_dim, n_model_size)
        self.encoder = MultiStageEncoder(n_model_size, **encoder_config)
        self.quantizer = MultiStageQuantizer(n_model_size, encoder_config["downsample_scales"][::-1],
                                             **quantizer_config)
        decoder_config["num_mels"] = n_model_size
        self.decoder = HifiGANGenerator(**decoder_config)
        if frame_decoder_config is not None:
            self.frame_decoder = FFTBlocks(d_model=n_model_size, name="frame_decoder", **frame_decoder_config)
        if pred_mel:
            self.mel_predictor = Ly.Linear(n_model_size, in_dim)

    def _decode_frames(self, decoder_inputs, length):
        if hasattr(self, "frame_decoder"):
            decoder_inputs, _ = self.frame_decoder(decoder_inputs, _positions(length, decoder_inputs.shape[1]))
        return decoder_inputs

    def _vocode(self, frames):
        """frames (B, T, C) -> waveform (B, S, 1); the generator consumes channels-last directly"""
        return self.decoder.forward_cl(frames)

    def forward(self, mel, mel_length, warmup=False, window=None):
        output_dict = {}
        encoder_states = self.encoder(self.in_linear(mel), mel_length)
        quantizer_states = self.quantizer(encoder_states)
        decoder_inputs = quantizer_states["residual_output"]
        encoder_outputs, encoder_lengths = zip(*encoder_states)
        output_dict.update({
            "encoder_outputs": encoder_outputs[::-1], "encoder_lengths": encoder_lengths[::-1],
            "encoder_indices": quantizer_states["quantizer_indices"],
            "encoder_diffs": quantizer_states["quantizer_diffs"],
            "decoder_diffs": quantizer_states["predictor_diffs"]})
        decoder_inputs = self._decode_frames(decoder_inputs, mel_length)
        if hasattr(self, "mel_predictor"):
            output_dict["mel_outputs"] = self.mel_predictor(decoder_inputs)
        if not warmup:
            if window is not None:
                if isinstance(window, tuple) and torch.is_tensor(window[0]):
                    # (starts (B,) device int64, n_frames): sync-free form used by the trainer
                    starts, n = window
                    idx = starts.view(-1, 1, 1) + torch.arange(n, device=starts.device).view(1, -1, 1)
                    decoder_inputs = torch.gather(decoder_inputs, 1, idx.expand(-1, -1, decoder_inputs.shape[2]))
                else:
                    assert len(window) == decoder_inputs.shape[0]
                    decoder_inputs = torch.stack([decoder_inputs[i, s:e] for i, (s, e) in enumerate(window)], dim=0)
            output_dict["decoder_outputs"] = self._vocode(decoder_inputs)
        return output_dict

    def analysis(self, mel, mel_length):
        encoder_states = self.encoder(self.in_linear(mel), mel_length)
        quantizer_states = self.quantizer(encoder_states)
        if self.training:
            encoder_outputs, encoder_lengths = zip(*encoder_states)
            return {"encoder_outputs": encoder_outputs[::-1], "encoder_lengths": encoder_lengths[::-1],
                    "encoder_indices": quantizer_states["quantizer_indices"],
                    "encoder_diffs": quantizer_states["quantizer_diffs"],
                    "decoder_diffs": quantizer_states["predictor_diffs"], "quantizer_states": quantizer_states}
        return quantizer_states

    def synthesis(self, quantizer_outputs, quantizer_lengths):
        quantizer_states = quantizer_outputs
        if not isinstance(quantizer_outputs, dict):
            quantizer_states = self.quantizer(zip(quantizer_outputs, quantizer_lengths), from_encoder=False)
        decoder_inputs = self._decode_frames(quantizer_states["residual_output"], quantizer_lengths[-1])
        decoder_outputs = self._vocode(decoder_inputs)
        if self.training:
            output_dict = {"decoder_outputs": decoder_outputs}
            if hasattr(self, "mel_predictor"):
                output_dict["mel_outputs"] = self.mel_predictor(decoder_inputs)
            return output_dict
        return decoder_outputs

    def compute_embedding_loss(self, quantizer_outputs, quantizer_lengths, quantizer_states, methods=["mse"],
                               loss_weights=[1.0]):
        pred_states = [{"predictor_outputs": quantizer_outputs[i],
                        "target_outputs": quantizer_states["quantizer_outputs"][i],
                        "target_indices": quantizer_states["quantizer_indices"][i],
                        "target_lengths": quantizer_lengths[i]} for i in range(len(quantizer_outputs))]
        return self.quantizer.compute_embedding_loss(pred_states, methods, loss_weights)
