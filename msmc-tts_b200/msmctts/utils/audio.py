"""Spectral front end of the multi-resolution discriminator on the sm_100a kernels
(reference utils/audio.py: create_fb_matrix :30-84, MelScale :314-376, TorchSTFT :379-426).

The windowed, normalised, centre-padded (reflect) STFT is one msmc_conv_forward over the raw waveform with a fixed
(n_fft x 2F) cos|-sin basis (Cs = 1, stride = hop); magnitude, mel projection and the 'double' lin/log stacking
follow as point-wise kernels + one GEMM.  Output layout is (B, frames, F, channels) -- the reference's
(B, channels, F, frames) with the spatial axes exchanged; `transform` returns the reference layout for API parity.
"""
import math

import numpy as np
import torch
from torch import nn

from msmctts._b200 import functional as Fn


def create_fb_matrix(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None):
    """HTK triangular filterbank clamped to [1e-6, 1] (reference audio.py:30-84, norm=None)"""
    if norm is not None:
        raise ValueError("only norm=None is used on this path")
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:])
    return torch.clamp(fb, 1e-6, 1)


def dft_basis(n_fft, win_length, normalized, dtype=torch.float32):
    """(win_length, 2*Fp) basis [cos | 0 | -sin | 0] * hann(win_length), for the window's non-zero span only; each
    half is zero-padded from F = n_fft/2+1 to Fp = a multiple of 32 so spectrum and magnitude are tensor-core
    GEMM eligibility).  torch.stft centres a short window inside n_fft (left pad (n_fft - win)/2).
    Returns (basis, left, Fp)."""
    left = (n_fft - win_length) // 2
    F = n_fft // 2 + 1
    Fp = (F + 31) // 32 * 32
    k = np.arange(win_length, dtype=np.float64) + left
    f = np.arange(F, dtype=np.float64)
    ang = 2.0 * np.pi * np.outer(k, f) / n_fft
    win = torch.hann_window(win_length, dtype=torch.float64).numpy()     # periodic, like the reference
    scale = (1.0 / math.sqrt(n_fft)) if normalized else 1.0
    win_p = (win_length + 31) // 32 * 32           # zero rows: the unfolded-frame GEMM wants K % 32 == 0
    basis = np.zeros((win_p, 2 * Fp))
    basis[:win_length, :F] = np.cos(ang) * win[:, None] * scale
    basis[:win_length, Fp:Fp + F] = -np.sin(ang) * win[:, None] * scale
    return torch.from_numpy(basis).to(dtype), left, Fp


class MelScale(nn.Module):
    def __init__(self, n_mels=128, sample_rate=24000, f_min=0.0, f_max=None, n_stft=None):
        super().__init__()
        self.n_mels, self.sample_rate = n_mels, sample_rate
        self.f_max = f_max if f_max is not None else float(sample_rate // 2)
        self.f_min = f_min
        fb = create_fb_matrix(n_stft, self.f_min, self.f_max, n_mels, sample_rate) if n_stft else torch.empty(0)
        self.register_buffer("fb", fb, persistent=False)
        # rows zero-padded to the padded spectrum width: the magnitude keeps its 32-aligned pitch and the filterbank
        # product is a tensor-core GEMM (pad columns of the magnitude meet zero weights)
        n_pad = (fb.shape[0] + 31) // 32 * 32 if n_stft else 0
        fb_p = torch.zeros(n_pad, n_mels)
        if n_stft:
            fb_p[:fb.shape[0]] = fb
        self.register_buffer("fb_p", fb_p, persistent=False)

    def forward_cl(self, mag):
        """mag (B, frames, F | Fp) -> (B, frames, n_mels)"""
        B, T, Fq = mag.shape
        fb = self.fb_p if Fq == self.fb_p.shape[0] else self.fb
        y = Fn.conv_cl(mag.reshape(B * T, 1, 1, Fq), fb, out_channels=self.n_mels)
        return y.reshape(B, T, self.n_mels)


class TorchSTFT(nn.Module):
    def __init__(self, fft_size, hop_size, win_size, normalized=False, domain="linear", mel_scale=False,
                 sample_rate=24000, ref_level_db=20, min_level_db=-100):
        super().__init__()
        self.fft_size, self.hop_size, self.win_size = fft_size, hop_size, win_size
        self.ref_level_db, self.min_level_db = ref_level_db, min_level_db
        self.normalized, self.domain = normalized, domain
        self.n_freq = fft_size // 2 + 1
        basis, self.left, self.n_freq_pad = dft_basis(fft_size, win_size, normalized)
        self.register_buffer("basis", basis, persistent=False)          # (win, 2Fp): GEMM layout [tap][1][2Fp]
        self.register_buffer("basis_t", basis[:win_size].t().contiguous(), persistent=False)
        self.mel_scale = MelScale(self.n_freq, sample_rate, n_stft=self.n_freq) if mel_scale else None

    def spectrum_cl(self, x, center=True, pad=None):
        """x (B, L) -> (B, frames, 2Fp) = [re | pad | im | pad]"""
        B, L = x.shape
        p = (self.fft_size // 2 if center else 0) if pad is None else pad
        p = p - self.left
        return Fn.stft_frames(x, self.basis, self.basis_t, self.hop_size, p, self.win_size)

    def transform_cl(self, x):
        """x (B, L) -> (B, frames, F, C) with C = 2 ('double': lin, log), else 1"""
        if self.mel_scale is not None:
            mag = self.mel_scale.forward_cl(Fn.spec_magnitude(self.spectrum_cl(x), 1e-7, False))   # padded width
        else:
            mag = Fn.spec_magnitude(self.spectrum_cl(x), 1e-7, False, self.n_freq)
        if self.domain == "double":
            return Fn.mel_double(mag, self.ref_level_db, self.min_level_db)
        if self.domain == "linear":
            return mag.unsqueeze(-1)
        return Fn.mel_double(mag, self.ref_level_db, self.min_level_db)[..., 1:]

    def transform(self, x):
        """reference layout: (B, 2F | F, frames), phase is not computed on this path (unused by the discriminator)"""
        y = self.transform_cl(x)                      # (B, frames, F, C)
        B, T, Fq, Cn = y.shape
        return y.permute(0, 3, 2, 1).reshape(B, Cn * Fq, T), None


class MelExtractor(nn.Module):
    """On-GPU log-mel features in the reference's normalisation (reference
    examples/csmsc/scripts/audio/audio.py:59-63,114-131 + hparams.py): pre-emphasis 0.97, STFT(n_fft, hop, win,
    hann, centre / reflect), Slaney mel filterbank (librosa.filters.mel, htk=False -- restated, see
    trainers/criterions/stft_loss.py), 20 log10(max(1e-5, .)) - ref_level_db, symmetric normalisation to
    [-max_abs, max_abs].  Replaces the librosa / scipy preprocessing pass: waveforms go to the device once and the
    features never touch the host.  wav (B, L) -> mel (B, frames, n_mels), frames = L // hop + 1."""

    def __init__(self, sample_rate=24000, n_fft=2048, hop_size=300, win_size=1200, n_mels=80, preemphasis=0.97,
                 ref_level_db=20.0, min_level_db=-100.0, max_abs_value=4.0):
        super().__init__()
        from msmctts.trainers.criterions.stft_loss import mel_filterbank_slaney
        self.n_fft, self.hop_size, self.win_size, self.n_mels = n_fft, hop_size, win_size, n_mels
        self.preemphasis, self.ref_level_db = preemphasis, ref_level_db
        self.min_level_db, self.max_abs_value = min_level_db, max_abs_value
        self.n_freq = n_fft // 2 + 1
        basis, self.left, self.n_freq_pad = dft_basis(n_fft, win_size, normalized=False)
        self.register_buffer("basis", basis, persistent=False)
        self.register_buffer("basis_t", basis[:win_size].t().contiguous(), persistent=False)
        mel_g = torch.zeros(self.n_freq_pad, n_mels)
        mel_g[:self.n_freq] = mel_filterbank_slaney(sample_rate, n_fft, n_mels, 0, sample_rate // 2).t()
        self.register_buffer("mel_basis_g", mel_g, persistent=False)

    @torch.no_grad()
    def forward(self, wav):
        if wav.dim() == 3:
            wav = wav.squeeze(-1)
        # y[n] = x[n] - a x[n-1]  (scipy.signal.lfilter([1, -a], [1], x), zero initial state)
        y = wav - self.preemphasis * torch.nn.functional.pad(wav, (1, 0))[:, :-1]
        spec = Fn.stft_frames(y.contiguous(), self.basis, self.basis_t, self.hop_size, self.n_fft // 2 - self.left,
                              self.win_size)
        mag = Fn.spec_magnitude(spec, 0.0, False)                       # |STFT|, padded width
        B, T, Fp = mag.shape
        mel = Fn.conv_cl(mag.reshape(B * T, 1, 1, Fp), self.mel_basis_g, out_channels=self.n_mels)
        db = 20.0 * torch.log10(mel.reshape(B, T, self.n_mels).clamp_min(1e-5)) - self.ref_level_db
        s = (2 * self.max_abs_value) * ((db - self.min_level_db) / (-self.min_level_db)) - self.max_abs_value
        return s.clamp(-self.max_abs_value, self.max_abs_value)
