"""CPU oracle of the mel front-end.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic lives in librosa (0.9.2, assets/environment.yml:54), which is absent from /root/reference and from
this image, and the reference holds no mel fixture.  This file restates librosa's published algorithm for the one call the
reference makes -- ``librosa.feature.melspectrogram(y=aud, sr=18000, hop_length=1200, n_mels=128)`` (trainers/ddpm_show_trainer.py:
1063, trainers/ddpm_beat_trainer.py:1244, datasets/beat.py:371) -- operation by operation:

* ``librosa.stft``: ``get_window('hann', 2048, fftbins=True)`` (float64), ``np.pad(y, 1024, mode=pad_mode)``, frames every hop,
  ``np.fft.rfft(window * frames)`` in float64, stored as complex64;
* ``_spectrogram``: ``np.abs(S) ** 2`` (float32);
* ``librosa.filters.mel``: Slaney scale, ``norm='slaney'`` triangles (float32), ``np.dot(mel_basis, S)``.
It is anchored instead (tests/test_wave_frontend.py) on two independent implementations -- torch.stft in float64 for framing /
padding / FFT, and transformers.audio_utils (Hugging Face's librosa-compatible window, Slaney filterbank and mel spectrogram:
1.7e-7 on the whole spectrogram, 2e-9 on the filterbank) -- and on properties: Parseval per frame, a pure tone lands in the band that
contains it, the filterbank's closed form (unit-area triangles, band edges).
"""
import numpy as np


def hann_window(n):
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def hz_to_mel(f):
    f = float(f)
    if f >= 1000.0:
        return 15.0 + np.log(f / 1000.0) / (np.log(6.4) / 27.0)
    return f / (200.0 / 3)


def mel_to_hz(m):
    m = float(m)
    if m >= 15.0:
        return 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    return (200.0 / 3) * m


def mel_basis(sr, n_fft, n_mels):
    freqs = np.arange(1 + n_fft // 2, dtype=np.float64) * (float(sr) / n_fft)
    mels = np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2)
    edges = np.array([mel_to_hz(m) for m in mels])
    W = np.zeros((n_mels, freqs.size), dtype=np.float64)
    for i in range(n_mels):
        lo, c, hi = edges[i], edges[i + 1], edges[i + 2]
        up = (freqs - lo) / (c - lo)
        down = (hi - freqs) / (hi - c)
        W[i] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (hi - lo))
    return W.astype(np.float32)


def melspectrogram(y, sr=18000, n_fft=2048, hop_length=1200, n_mels=128, pad_mode="constant"):
    """-> float32 [n_mels, 1 + len(y) // hop_length] like librosa (frames along the LAST axis)."""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (yp.size - n_fft) // hop_length
    win = hann_window(n_fft)
    S = np.empty((1 + n_fft // 2, n_frames), dtype=np.complex64)
    for f in range(n_frames):
        S[:, f] = np.fft.rfft(win * yp[f * hop_length:f * hop_length + n_fft].astype(np.float64))
    power = (np.abs(S) ** 2).astype(np.float32)
    return np.dot(mel_basis(sr, n_fft, n_mels), power)
