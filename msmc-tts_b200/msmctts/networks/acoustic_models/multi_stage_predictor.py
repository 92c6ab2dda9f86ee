"""Multi-stage predictor (text -> per-stage quantised-feature predictions) on the sm_100a kernels.
Same class / kwargs / state_dict keys as reference acoustic_models/multi_stage_predictor.py:9-126."""
import torch
import torch.nn.functional as F
from torch import nn

from msmctts._b200 import layers as Ly
from .transformer import FFTBlocks, LengthRegulator


def _positions(lengths, t):
    ids = torch.arange(1, t + 1, device=lengths.device).unsqueeze(0)
    return torch.where(ids <= lengths.unsqueeze(1), ids, torch.zeros_like(ids)).long()


class _Decoder(nn.ModuleList):
    pass


class MultiStagePredictor(nn.Module):
    def __init__(self, n_symbols, n_model_size, n_pred_size, n_pred_scale, encoder_config, adaptor_config,
                 decoder_config):
        super().__init__()
        self.n_pred_scale = list(n_pred_scale)
        self.n_symbols = n_symbols
        if isinstance(n_symbols, (tuple, list)):
            self.word_emb = nn.ModuleList([nn.Embedding(n, n_model_size, padding_idx=0) for n in n_symbols])
        else:
            self.word_emb = nn.Embedding(n_symbols, n_model_size, padding_idx=0)
        self.encoder = FFTBlocks(**encoder_config)
        self.upsampler = LengthRegulator(**adaptor_config)
        self.downsamplers = nn.ModuleList([
            Ly.Conv1d(n_model_size, n_model_size, scale * 2 + 1, padding=scale) for scale in self.n_pred_scale[::-1]])
        self.decoders = nn.ModuleList([
            nn.ModuleList([Ly.Linear(n_model_size * 2 + n_pred_size if i > 0 else n_model_size, n_model_size),
                           FFTBlocks(**decoder_config), Ly.Linear(n_model_size, n_pred_size)])
            for i in range(len(self.n_pred_scale))])

    def forward(self, text, text_length, dur=None, feat=None, feat_length=None):
        output, duration = self.encode(text, text_length, dur)
        if feat_length is None:
            total = duration.sum(-1).long()
            feat_length = []
            for scale in self.n_pred_scale[::-1]:
                total = torch.ceil(total / scale).long()
                feat_length.append(total)
            feat_length = feat_length[::-1]
        output = self.decode(output, feat, feat_length)
        return {"feat": output, "feat_length": feat_length, "text_length": text_length, "duration": duration}

    def encode(self, text, text_length, dur=None):
        if isinstance(self.n_symbols, (tuple, list)):
            output = sum(self.word_emb[i](text[..., i].long()) for i in range(len(self.word_emb)))
        else:
            output = self.word_emb(text.long())
        output, text_mask = self.encoder(output, _positions(text_length, text.shape[1]))
        output, _, duration = self.upsampler(output, text_mask, target=dur, alpha=1.0)
        return output, duration

    def decode(self, text_embedding, feat=None, feat_lengths=None):
        downsampled = []
        for model, scale in zip(self.downsamplers, self.n_pred_scale[::-1]):
            text_embedding = model(text_embedding)
            text_embedding = F.avg_pool1d(text_embedding.transpose(1, 2), kernel_size=scale, stride=scale,
                                          ceil_mode=True).transpose(1, 2)
            downsampled.append(text_embedding)
        downsampled = downsampled[::-1]
        predictions = []
        output = None
        for i, decoder in enumerate(self.decoders):
            text_embedding = downsampled[i]
            pos = _positions(feat_lengths[i], text_embedding.shape[1])
            if i > 0:
                scale = self.n_pred_scale[i - 1]
                pre_input = feat[i - 1] if feat is not None else predictions[-1]
                pre_input = torch.cat((output, pre_input), dim=2)
                pre_input = torch.repeat_interleave(pre_input, scale, dim=1)[:, : text_embedding.shape[1]]
                output = torch.cat((text_embedding, pre_input), dim=2)
            else:
                output = text_embedding
            output = decoder[0](output)
            output, _ = decoder[1](output, pos)
            prediction = decoder[2](output)
            if not self.training and hasattr(self, "quantizers"):
                q = self.quantizers[i]
                prediction = (q.quantize if hasattr(q, "quantize") else q)(prediction)[0]
            predictions.append(prediction)
        return predictions
