"""Import the UNMODIFIED reference (hhguo/MSMC-TTS, /root/reference) for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py in the build container
(the reference tree does not exist on the GPU box, nothing at run time may touch it).

Shim list follows SURVEY.md section 8(c); every shim is inert w.r.t. arithmetic except the
labelled `torch.stft` wrapper (the reference calls torch.stft without return_complex,
utils/audio.py:399-402, criterions/stft_loss.py:21-23,95-102, which torch>=2.0 rejects).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("MSMC_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """Restatement of librosa.filters.mel (htk=False, norm='slaney'), librosa>=0.8 published algorithm.

    librosa is an un-vendored third-party dependency (requirements.txt:6, unpinned) that is not
    installable here: PARITY UNPINNED for this filterbank.
    """
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        freqs = f_sp * m
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)

    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


_installed = False


def install():
    """Install shims and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not found at %s (golden generation runs in the build container only)" % REF_ROOT)
    _stub("turtle", update=None)                               # B1: msmc_vqgan.py:1
    _stub("soundfile", SoundFile=object, read=None, write=None)  # utils/utils.py:2,13
    _stub("tensorboardX", SummaryWriter=object)                # utils/logger.py:1
    lib = _stub("librosa")
    lib.util = _stub("librosa.util", pad_center=None, tiny=None, normalize=None)
    lib.filters = _stub("librosa.filters",
                        mel=lambda sr, n_fft, n_mels, fmin, fmax: slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax))
    # Labelled NON-INERT shim: legacy real-view stft semantics on torch>=2.
    _orig_stft = torch.stft

    def _stft_compat(input, n_fft, hop_length=None, win_length=None, window=None, center=True,
                     pad_mode="reflect", normalized=False, onesided=None, return_complex=None):
        out = _orig_stft(input, n_fft, hop_length, win_length, window, center, pad_mode, normalized,
                         onesided, return_complex=True)
        return torch.view_as_real(out) if return_complex is None else out

    torch.stft = _stft_compat
    sys.path.insert(0, REF_ROOT)
    # B2: vqgantts/__init__.py imports a file that is not in the tree; pre-register a bare package.
    import msmctts.networks  # noqa: F401
    pkg = types.ModuleType("msmctts.networks.vqgantts")
    pkg.__path__ = [os.path.join(REF_ROOT, "msmctts", "networks", "vqgantts")]
    sys.modules["msmctts.networks.vqgantts"] = pkg
    _installed = True


def ref(modname):
    install()
    return importlib.import_module(modname)
