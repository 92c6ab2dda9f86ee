"""MelLoss on the sm_100a kernels (reference trainers/criterions/stft_loss.py:55-114).

reflect-pad (fft-hop)/2, STFT(fft, hop, hann(win), center=False), sqrt(re^2+im^2+1e-9), Slaney mel filterbank,
log(clamp(., 1e-5)), L1.  The framing + windowed DFT is one msmc_conv_forward over the raw waveform restricted to the
window's non-zero span (win taps instead of fft taps).  The reference takes the filterbank from
librosa.filters.mel (third-party, unpinned, not vendored): `mel_filterbank_slaney` restates librosa's published
algorithm -- PARITY UNPINNED for that matrix; everything else is pinned by tests/golden/melloss.pt.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from msmctts._b200 import functional as Fn
from msmctts.utils.audio import dft_basis


def mel_filterbank_slaney(sr, n_fft, n_mels, fmin, fmax):
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    pts = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    ramps = pts[:, None] - freqs[None, :]
    fdiff = np.diff(pts)
    w = np.maximum(0.0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
    w *= (2.0 / (pts[2:n_mels + 2] - pts[:n_mels]))[:, None]
    return torch.from_numpy(w.astype(np.float32))


class MelLoss(nn.Module):
    def __init__(self, fft_size, hop_size, win_size, sample_rate, num_mels):
        super().__init__()
        self.fft_size, self.hop_size, self.win_size = fft_size, hop_size, win_size
        self.sample_rate, self.num_mels = sample_rate, num_mels
        self.n_freq = fft_size // 2 + 1
        basis, self.left, self.n_freq_pad = dft_basis(fft_size, win_size, normalized=False)
        self.register_buffer("basis", basis, persistent=False)
        self.register_buffer("basis_t", basis[:win_size].t().contiguous(), persistent=False)
        mel_basis = mel_filterbank_slaney(sample_rate, fft_size, num_mels, 0, sample_rate // 2)   # (num_mels, F)
        self.register_buffer("mel_basis", mel_basis, persistent=False)
        # GEMM layout (Fp, num_mels), rows past F zero: padded magnitude x filterbank on the tensor cores
        mel_g = torch.zeros(self.n_freq_pad, num_mels)
        mel_g[:self.n_freq] = mel_basis.t()
        self.register_buffer("mel_basis_g", mel_g, persistent=False)

    def mel_spectrogram(self, y):
        """y (B, L) -> log-mel (B, frames, num_mels)"""
        B, L = y.shape
        pad = int((self.fft_size - self.hop_size) / 2) - self.left
        spec = Fn.stft_frames(y, self.basis, self.basis_t, self.hop_size, pad, self.win_size)
        mag = Fn.spec_magnitude(spec, 1e-9, True)                       # (B, frames, Fp), pad columns meet zero weights
        Bf, Tf, Fp = mag.shape
        mel = Fn.conv_cl(mag.reshape(Bf * Tf, 1, 1, Fp), self.mel_basis_g,
                         out_channels=self.num_mels).reshape(Bf, Tf, self.num_mels)
        return Fn.log_clamp(mel, 1e-5)

    def forward(self, predicts, targets):
        with torch.no_grad():
            t = self.mel_spectrogram(targets)
        return F.l1_loss(self.mel_spectrogram(predicts), t)
