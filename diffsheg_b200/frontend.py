"""Audio front-end without network weights: the mel spectrogram of the trainers' inference path on the GPU.

    mel = librosa.feature.melspectrogram(y=aud, sr=18000, hop_length=1200, n_mels=128)     # show:1063, beat:1244, datasets/beat.py:371
    mel = mel[..., :-1]; audio_emb = torch.from_numpy(np.swapaxes(mel, -1, -2)).unsqueeze(0)  # show:1065-1067

librosa (0.9.2 in assets/environment.yml:54) is not a dependency of this package: the window and the Slaney filterbank it would
build are restated here from its published algorithm (``scipy.signal.get_window('hann', n, fftbins=True)``;
``librosa.filters.mel(htk=False, norm='slaney')``) in float64 numpy, uploaded once per device, and the framing / FFT / power /
filterbank product run in csrc/frontend.cuh.  HuBERT extraction -- the other half of f1 -- needs downloaded weights and stays out.
"""
import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream

N_FFT = 2048
PAD_MODES = {"constant": 0, "reflect": 1}


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True) = general_cosine(n + 1, [0.5, 0.5])[:-1] (float64)."""
    fac = np.linspace(-np.pi, np.pi, n + 1)
    return (0.5 + 0.5 * np.cos(fac))[:-1]


def _hz_to_mel(f):
    """Slaney scale (librosa.hz_to_mel, htk=False): linear below 1 kHz (200/3 Hz per mel), logarithmic above (27 mels per factor 6.4)."""
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels, htk=False, norm='slaney') -> float32 [n_mels, 1 + n_fft // 2]:
    triangles between n_mels + 2 band edges equally spaced on the Slaney mel scale, each scaled by 2 / (its width in Hz)."""
    fmax = float(sr) / 2 if fmax is None else fmax
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(edges)
    ramps = edges[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    weights *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[:, None]
    return weights.astype(np.float32)


def band_ranges(basis):
    """[n_mels, 2] int32: first and one-past-last non-zero bin of every band (an empty band gets 0, 0)."""
    nz = basis != 0
    lo = np.where(nz.any(1), nz.argmax(1), 0)
    hi = np.where(nz.any(1), basis.shape[1] - nz[:, ::-1].argmax(1), 0)
    return np.stack([lo, hi], 1).astype(np.int32)


_tables = {}


def _device_tables(device, sr, n_mels):
    key = (device.index, int(sr), int(n_mels))
    if key not in _tables:
        basis = mel_filterbank(sr, N_FFT, n_mels)
        _tables[key] = (torch.from_numpy(hann_periodic(N_FFT).astype(np.float32)).to(device),
                        torch.from_numpy(basis).to(device), torch.from_numpy(band_ranges(basis)).to(device))
    return _tables[key]


def mel_spectrogram(audio, sr=18000, hop_length=1200, n_mels=128, pad_mode="constant", drop_last=False):
    """``librosa.feature.melspectrogram(y=audio, sr=sr, hop_length=hop_length, n_mels=n_mels)`` for a CUDA fp32 waveform ``[n]``,
    returned frame-major ``[n_frames, n_mels]`` (= ``np.swapaxes(mel, -1, -2)``, show:1066), ``n_frames = 1 + n // hop_length``.
    ``pad_mode``: librosa.stft's centring pad ('constant' zeros or 'reflect').  ``drop_last``: the trainers' ``mel[..., :-1]`` (show:1065)."""
    if not (torch.is_tensor(audio) and audio.is_cuda and audio.dim() == 1):
        raise ValueError("mel_spectrogram expects a 1-D CUDA tensor")
    if pad_mode not in PAD_MODES:
        raise ValueError(f"pad_mode must be one of {sorted(PAD_MODES)}")
    x = audio.to(torch.float32).contiguous()
    n = x.numel()
    n_frames = 1 + n // int(hop_length) - (1 if drop_last else 0)
    if n_frames < 1:
        raise ValueError("audio too short for one frame")
    window, basis, ranges = _device_tables(x.device, sr, n_mels)
    out = torch.empty(n_frames, int(n_mels), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dsheg_mel_spectrogram(_ptr(x), n, N_FFT, int(hop_length), PAD_MODES[pad_mode], _ptr(window), _ptr(basis),
                                                    _ptr(ranges), int(n_mels), _ptr(out), n_frames, _stream(x.device)), None, "mel_spectrogram")
    return out


def audio_embedding(audio_18k, pad_mode="constant"):
    """show:1063-1067 / beat:1244-1248: the ``audio_emb`` tensor ``[1, N, 128]`` the window loop slices, from the 18 kHz waveform."""
    return mel_spectrogram(audio_18k, 18000, 1200, 128, pad_mode, drop_last=True).unsqueeze(0)
