"""Mel front-end (trainers/ddpm_show_trainer.py:1063-1067): oracle anchoring, host tables, the kernel source on the
CPU emulator, and -- on a GPU -- the product through the C ABI.

librosa is absent (no network), so the oracle (oracle/frontend.py) is a restatement of its published algorithm: PARITY UNPINNED
against librosa itself.  It is anchored here on an independent STFT (torch.stft in float64), on Parseval's identity and on the closed
form of the Slaney filterbank; the CUDA kernel is then held to the oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

from diffsheg_b200 import frontend as fe
from oracle import frontend as ofe

SR, HOP, N_MELS = 18000, 1200, 128


def _audio(n, seed=0):
    """speech-like test signal: three tones with a slow envelope + noise, amplitude ~0.3"""
    g = np.random.default_rng(seed)
    t = np.arange(n) / SR
    y = 0.2 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 1760 * t + 1.0) + 0.05 * np.sin(2 * np.pi * 6100 * t)
    return ((1 + 0.5 * np.sin(2 * np.pi * 1.3 * t)) * y + 0.02 * g.standard_normal(n)).astype(np.float32)


def _relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


# ---- the oracle and the host tables -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_oracle_agrees_with_an_independent_stft(pad_mode):
    y = _audio(18000 * 3 + 517, seed=1)
    want = ofe.melspectrogram(y, SR, 2048, HOP, N_MELS, pad_mode)
    win = torch.from_numpy(fe.hann_periodic(2048))
    S = torch.stft(torch.from_numpy(y).double(), 2048, HOP, window=win, center=True, pad_mode=pad_mode, return_complex=True)
    got = ofe.mel_basis(SR, 2048, N_MELS).astype(np.float64) @ (S.abs() ** 2).numpy()
    assert want.shape == got.shape == (N_MELS, 1 + y.size // HOP)
    assert _relmax(want, got) < 2e-6


def test_oracle_agrees_with_the_librosa_compatible_implementation_of_transformers():
    """transformers.audio_utils restates librosa too (its Slaney filterbank and STFT are the ones HF's feature extractors were
    validated against librosa with) -- a third-party implementation of the same published algorithm, present in this image:
    window, filterbank and the whole mel spectrogram must agree with the oracle and with the product's host tables."""
    au = pytest.importorskip("transformers.audio_utils")
    fb = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=N_MELS, min_frequency=0.0, max_frequency=SR / 2,
                            sampling_rate=SR, norm="slaney", mel_scale="slaney")
    assert np.abs(fb.T - fe.mel_filterbank(SR, 2048, N_MELS)).max() < 1e-8
    win = au.window_function(2048, "hann", periodic=True)
    assert np.abs(win - fe.hann_periodic(2048)).max() < 1e-12
    y = _audio(SR * 2 + 123, seed=7)
    got = au.spectrogram(y.astype(np.float64), win, frame_length=2048, hop_length=HOP, fft_length=2048, power=2.0, center=True,
                         pad_mode="constant", mel_filters=fb)
    want = ofe.melspectrogram(y, SR, 2048, HOP, N_MELS, "constant")
    assert got.shape == want.shape and _relmax(want, got) < 2e-6


def test_oracle_frames_satisfy_parseval_and_a_tone_lands_in_its_band():
    y = _audio(HOP * 6, seed=2)
    yp = np.pad(y, 1024)
    win = ofe.hann_window(2048)
    for f in (0, 3, 6):
        x = win * yp[f * HOP:f * HOP + 2048]
        X = np.fft.rfft(x)
        p = np.abs(X) ** 2
        assert abs((p[0] + p[-1] + 2 * p[1:-1].sum()) / 2048 - (x ** 2).sum()) < 1e-9 * max(1.0, (x ** 2).sum())
    tone = (0.5 * np.sin(2 * np.pi * 3000.0 * np.arange(HOP * 8) / SR)).astype(np.float32)
    mel = ofe.melspectrogram(tone, SR, 2048, HOP, N_MELS)
    edges = np.array([ofe.mel_to_hz(m) for m in np.linspace(0, ofe.hz_to_mel(SR / 2), N_MELS + 2)])
    band = int(np.argmax(mel[:, 4]))
    assert edges[band] < 3000.0 < edges[band + 2]


def test_filterbank_is_librosas_slaney_form():
    B = fe.mel_filterbank(SR, 2048, N_MELS)
    assert B.dtype == np.float32 and B.shape == (N_MELS, 1025)
    np.testing.assert_allclose(B, ofe.mel_basis(SR, 2048, N_MELS), rtol=0, atol=1e-9)     # two independent restatements
    assert (B >= 0).all() and (B.sum(1) > 0).all()                                       # no empty band at n_fft = 2048
    # Slaney scale: 200/3 Hz per mel below 1 kHz; 'slaney' norm: every triangle has unit area in Hz (up to bin sampling)
    assert abs(float(fe._mel_to_hz(3.0)) - 200.0) < 1e-9 and abs(float(fe._hz_to_mel(1000.0)) - 15.0) < 1e-12
    area = B.astype(np.float64).sum(1) * (SR / 2048)
    assert np.abs(area[20:] - 1.0).max() < 0.05
    r = fe.band_ranges(B)
    for m in range(N_MELS):
        assert (B[m, :r[m, 0]] == 0).all() and (B[m, r[m, 1]:] == 0).all() and B[m, r[m, 0]] > 0 and B[m, r[m, 1] - 1] > 0
    np.testing.assert_array_equal(fe.hann_periodic(2048), __import__("scipy.signal", fromlist=["get_window"]).get_window("hann", 2048, fftbins=True))


# ---- the kernel source on the CPU emulator ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,pad_mode", [(HOP * 3 + 100, "constant"), (HOP * 2 + 7, "reflect"), (700, "constant")])
def test_mel_kernel_source_on_emulator(n, pad_mode):
    import emu
    L = emu.lib()
    y = _audio(n, seed=3)
    win = fe.hann_periodic(2048).astype(np.float32)
    basis = fe.mel_filterbank(SR, 2048, N_MELS)
    rng = fe.band_ranges(basis)
    n_frames = 1 + n // HOP
    out = np.full((n_frames, N_MELS), np.nan, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = L.emu_mel_spectrogram(P(y), ctypes.c_longlong(n), HOP, fe.PAD_MODES[pad_mode], P(win), P(basis), P(rng), N_MELS, P(out), n_frames)
    assert rc == 0, L.emu_last_error().decode()
    want = ofe.melspectrogram(y, SR, 2048, HOP, N_MELS, pad_mode).T
    assert np.isfinite(out).all()
    assert _relmax(out, want) < 1e-5
    # schedule independence (the emulator's racecheck stand-in): a shuffled thread order must give the same bits -- a missing
    # __syncthreads between two butterfly stages would not
    out2 = np.full_like(out, np.nan)
    os.environ["EMU_SCHED"] = "shuffle"
    try:
        assert L.emu_mel_spectrogram(P(y), ctypes.c_longlong(n), HOP, fe.PAD_MODES[pad_mode], P(win), P(basis), P(rng), N_MELS, P(out2), n_frames) == 0
    finally:
        del os.environ["EMU_SCHED"]
    assert np.array_equal(out, out2)


def test_mel_entry_point_of_the_c_abi_on_the_emulated_library():
    """dsheg_mel_spectrogram itself (argument checks + launch code of csrc/engine.cu) through ctypes on the emulated library."""
    import emu
    L = emu.engine_lib()
    n = HOP * 2 + 300
    y = _audio(n, seed=6)
    win = fe.hann_periodic(2048).astype(np.float32)
    basis = fe.mel_filterbank(SR, 2048, N_MELS)
    rng = fe.band_ranges(basis)
    n_frames = 1 + n // HOP
    out = np.full((n_frames, N_MELS), np.nan, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    call = lambda n_fft, pad, frames, samples=n: L.dsheg_mel_spectrogram(P(y), samples, n_fft, HOP, pad, P(win), P(basis), P(rng), N_MELS, P(out), frames, None)   # noqa: E731
    assert call(2048, 0, n_frames) == 0, L.dsheg_last_error(None)
    assert _relmax(out, ofe.melspectrogram(y, SR, 2048, HOP, N_MELS).T) < 1e-5
    assert call(1024, 0, n_frames) != 0 and b"n_fft must be 2048" in L.dsheg_last_error(None)
    assert call(2048, 2, n_frames) != 0                      # unknown pad mode
    assert call(2048, 0, n_frames + 1) != 0                  # more frames than 1 + n_samples / hop
    assert call(2048, 1, 1, samples=900) != 0                # reflect padding needs more than n_fft / 2 samples


# ---- the product on a GPU ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("seconds,pad_mode", [(60.0, "constant"), (7.3, "reflect"), (0.2, "constant")])
def test_mel_spectrogram_matches_oracle(seconds, pad_mode):
    y = _audio(int(SR * seconds), seed=4)
    got = fe.mel_spectrogram(torch.from_numpy(y).cuda(), SR, HOP, N_MELS, pad_mode).cpu().numpy()
    want = ofe.melspectrogram(y, SR, 2048, HOP, N_MELS, pad_mode).T
    assert got.shape == want.shape == (1 + y.size // HOP, N_MELS)
    assert np.isfinite(got).all()
    assert _relmax(got, want) < 1e-5                         # fp32 butterflies vs float64 rfft rounded to complex64
    big = want > 1e-4 * want.max()                           # and band by band where a band holds signal
    assert float(np.abs(got[big] / want[big] - 1).max()) < 2e-3


@pytest.mark.gpu
def test_audio_embedding_is_the_trainers_tensor_and_bad_calls_fail_loudly():
    y = _audio(SR * 4, seed=5)
    emb = fe.audio_embedding(torch.from_numpy(y).cuda())
    want = np.swapaxes(ofe.melspectrogram(y, SR, 2048, HOP, N_MELS)[..., :-1], -1, -2)[None]      # show:1063-1067
    assert tuple(emb.shape) == want.shape == (1, y.size // HOP, N_MELS)
    assert _relmax(emb.cpu().numpy(), want) < 1e-5
    with pytest.raises(ValueError):
        fe.mel_spectrogram(torch.from_numpy(y), SR, HOP, N_MELS)                                   # host tensor: no CPU fallback
    with pytest.raises(RuntimeError):
        fe.mel_spectrogram(torch.from_numpy(y[:500]).cuda(), SR, HOP, N_MELS, "reflect")          # reflect needs > 1024 samples
