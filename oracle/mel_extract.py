"""CPU restatement (numpy) of the reference's offline mel extraction, examples/csmsc/scripts/audio/audio.py:59-63
(`melspectrogram`) with :24-25 (pre-emphasis), :73-75 (`_stft` -> librosa.stft), :96-101,114-116 (mel basis),
:118-131 (dB + symmetric normalisation).  TEST INFRASTRUCTURE ONLY.

librosa is third-party and not installable here (requirements.txt: librosa>=0.8.0, unpinned): `librosa.stft` is
restated from its documented behaviour at 0.8 -- centre-padded with pad_mode='reflect', periodic Hann window of
`win_length` zero-padded (centred) to `n_fft`, frames = 1 + len // hop -- and `librosa.filters.mel` (Slaney, htk=False)
as in oracle/ref_modules.py.  PARITY UNPINNED for those two third-party pieces; the rest follows the file above."""
import numpy as np

from .ref_modules import slaney_mel_filterbank


def melspectrogram(y, sample_rate=24000, n_fft=2048, hop=300, win=1200, n_mels=80, preemphasis=0.97,
                   ref_level_db=20.0, min_level_db=-100.0, max_abs_value=4.0):
    """y (L,) float -> (frames, n_mels) in [-max_abs_value, max_abs_value]"""
    y = np.asarray(y, dtype=np.float64)
    y = y - preemphasis * np.concatenate([[0.0], y[:-1]])                 # lfilter([1, -a], [1], y)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win) / win)            # scipy get_window('hann', fftbins=True)
    left = (n_fft - win) // 2
    window = np.zeros(n_fft)
    window[left:left + win] = w
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + len(y) // hop
    frames = np.stack([yp[i * hop: i * hop + n_fft] * window for i in range(n_frames)])
    mag = np.abs(np.fft.rfft(frames, n=n_fft, axis=1))                    # (frames, F)
    basis = slaney_mel_filterbank(sample_rate, n_fft, n_mels, 0, sample_rate // 2).numpy().astype(np.float64)
    mel = mag @ basis.T
    db = 20.0 * np.log10(np.maximum(1e-5, mel)) - ref_level_db
    s = (2 * max_abs_value) * ((db - min_level_db) / (-min_level_db)) - max_abs_value
    return np.clip(s, -max_abs_value, max_abs_value)
