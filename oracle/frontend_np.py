"""ORACLE (test infrastructure, never shipped or timed as the product): float64 numpy restatement of the
reference's log-mel front end.

Follows, function by function,
  torchvggish/mel_features.py:21-45   frame
  torchvggish/mel_features.py:48-68   periodic_hann
  torchvggish/mel_features.py:71-92   stft_magnitude
  torchvggish/mel_features.py:96-111  hertz_to_mel
  torchvggish/mel_features.py:114-189 spectrogram_to_mel_matrix
  torchvggish/mel_features.py:192-223 log_mel_spectrogram
  torchvggish/vggish_input.py:30-82   waveform_to_examples (16 kHz branch; resampling is third-party + unpinned)
with the constants of torchvggish/vggish_params.py:22-36.

Parity pin: the reference ships no tests or golden vectors.  This restatement is pinned against the
reference's own Python executed in the build container (tests/golden/make_golden.py imports /root/reference and
stores its outputs under tests/golden/); tests/test_oracle_golden.py checks the restatement against those
vectors bit for bit.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
STFT_WINDOW_SECONDS = 0.025
STFT_HOP_SECONDS = 0.010
NUM_MEL_BINS = 64
MEL_MIN_HZ = 125
MEL_MAX_HZ = 7500
LOG_OFFSET = 0.01
EXAMPLE_WINDOW_SECONDS = 0.96
EXAMPLE_HOP_SECONDS = 0.96


def num_frames(num_samples: int, window: int, hop: int) -> int:
    return 1 + int(np.floor((num_samples - window) / hop))


def frame(data: np.ndarray, window: int, hop: int) -> np.ndarray:
    """(num_frames, window, ...) copy of the overlapping frames; incomplete tail frames are dropped.
    A negative frame count raises ValueError like the reference's as_strided call does."""
    n = num_frames(data.shape[0], window, hop)
    if n < 0:
        raise ValueError("negative dimensions are not allowed")
    idx = hop * np.arange(n)[:, None] + np.arange(window)[None, :]
    return data[idx]


def periodic_hann(window: int) -> np.ndarray:
    return 0.5 - (0.5 * np.cos(2 * np.pi / window * np.arange(window)))


def stft_magnitude(signal: np.ndarray, fft_length: int, hop: int, window: int) -> np.ndarray:
    windowed = frame(signal, window, hop) * periodic_hann(window)
    return np.abs(np.fft.rfft(windowed, int(fft_length)))


def hertz_to_mel(hz):
    return 1127.0 * np.log(1.0 + (hz / 700.0))


def mel_matrix(num_mel_bins=NUM_MEL_BINS, num_spectrogram_bins=257, sample_rate=SAMPLE_RATE,
               lower_hz=MEL_MIN_HZ, upper_hz=MEL_MAX_HZ) -> np.ndarray:
    nyquist = sample_rate / 2.
    if lower_hz < 0.0:
        raise ValueError("lower_edge_hertz %.1f must be >= 0" % lower_hz)
    if lower_hz >= upper_hz:
        raise ValueError("lower_edge_hertz %.1f >= upper_edge_hertz %.1f" % (lower_hz, upper_hz))
    if upper_hz > nyquist:
        raise ValueError("upper_edge_hertz %.1f is greater than Nyquist %.1f" % (upper_hz, nyquist))
    bins_mel = hertz_to_mel(np.linspace(0.0, nyquist, num_spectrogram_bins))
    edges = np.linspace(hertz_to_mel(lower_hz), hertz_to_mel(upper_hz), num_mel_bins + 2)
    out = np.empty((num_spectrogram_bins, num_mel_bins))
    for band in range(num_mel_bins):
        lo, centre, hi = edges[band:band + 3]
        rising = (bins_mel - lo) / (centre - lo)
        falling = (hi - bins_mel) / (hi - centre)
        out[:, band] = np.maximum(0.0, np.minimum(rising, falling))
    out[0, :] = 0.0
    return out


def log_mel_spectrogram(data: np.ndarray) -> np.ndarray:
    """(num_frames, 64) float64 log-mel of a mono 16 kHz signal with the VGGish parameters."""
    window = int(round(SAMPLE_RATE * STFT_WINDOW_SECONDS))
    hop = int(round(SAMPLE_RATE * STFT_HOP_SECONDS))
    fft_length = 2 ** int(np.ceil(np.log(window) / np.log(2.0)))
    spec = stft_magnitude(data, fft_length, hop, window)
    mel = np.dot(spec, mel_matrix(num_spectrogram_bins=spec.shape[1]))
    return np.log(mel + LOG_OFFSET)


def waveform_to_examples(data: np.ndarray, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """(num_examples, 96, 64) float64.  `data` is (samples,) or (samples, channels) as the reference's
    np.mean(axis=1) expects (vggish_input.py:49-50)."""
    if len(data.shape) > 1:
        data = np.mean(data, axis=1)
    if sample_rate != SAMPLE_RATE:
        raise NotImplementedError("resampling (resampy) is outside the pinned path")
    log_mel = log_mel_spectrogram(data)
    rate = 1.0 / STFT_HOP_SECONDS
    win = int(round(EXAMPLE_WINDOW_SECONDS * rate))
    hop = int(round(EXAMPLE_HOP_SECONDS * rate))
    return frame(log_mel, win, hop)
