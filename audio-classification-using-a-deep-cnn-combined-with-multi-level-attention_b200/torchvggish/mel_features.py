"""Drop-in for torchvggish/mel_features.py.

The reference computes everything here in float64 numpy (mel_features.py:21-223).  In this build the heavy
part — framing, Hann windowing, the 512-point real DFT magnitude, the mel projection and the log — is ONE CUDA
kernel behind `b200.engine.logmel` (csrc/frontend.cu); the functions below keep the reference's names, arguments
and error behaviour on top of it.  The kernel is specialised for the VGGish parameters (16 kHz, 25 ms / 10 ms,
64 bands 125-7500 Hz); other parameter sets raise NotImplementedError rather than silently running on the CPU.
"""
import numpy as np
import torch

from b200 import engine as _engine

_MEL_BREAK_FREQUENCY_HERTZ = 700.0
_MEL_HIGH_FREQUENCY_Q = 1127.0


def frame(data, window_length, hop_length):
    """(num_frames, window_length, ...) view of successive frames, hop_length apart; the incomplete tail is
    dropped and nothing is padded (reference mel_features.py:21-45).  Works on numpy arrays and torch tensors
    (any device); pure indexing, no arithmetic."""
    num_samples = data.shape[0]
    num_frames = 1 + int(np.floor((num_samples - window_length) / hop_length))
    if num_frames < 0:
        raise ValueError("negative dimensions are not allowed")
    if isinstance(data, torch.Tensor):
        size = (num_frames, window_length) + tuple(data.shape[1:])
        stride = (data.stride(0) * hop_length,) + tuple(data.stride())
        return data.as_strided(size, stride)
    shape = (num_frames, window_length) + data.shape[1:]
    strides = (data.strides[0] * hop_length,) + data.strides
    return np.lib.stride_tricks.as_strided(data, shape=shape, strides=strides)


def periodic_hann(window_length):
    """Periodic Hann window (reference mel_features.py:48-68).  For the VGGish length (400) this is the very
    table the kernel folds into its DFT basis."""
    if int(window_length) == 400:
        return _engine.front_end_tables()[0]
    return 0.5 - (0.5 * np.cos(2 * np.pi / window_length * np.arange(window_length)))


def hertz_to_mel(frequencies_hertz):
    """HTK mel scale (reference mel_features.py:100-111)."""
    return _MEL_HIGH_FREQUENCY_Q * np.log(1.0 + (frequencies_hertz / _MEL_BREAK_FREQUENCY_HERTZ))


def spectrogram_to_mel_matrix(num_mel_bins=20, num_spectrogram_bins=129, audio_sample_rate=8000,
                              lower_edge_hertz=125.0, upper_edge_hertz=3800.0):
    """(num_spectrogram_bins, num_mel_bins) HTK mel weights (reference mel_features.py:114-189), including its
    ValueErrors.  For the VGGish configuration the library's own table is returned."""
    nyquist_hertz = audio_sample_rate / 2.
    if lower_edge_hertz < 0.0:
        raise ValueError("lower_edge_hertz %.1f must be >= 0" % lower_edge_hertz)
    if lower_edge_hertz >= upper_edge_hertz:
        raise ValueError("lower_edge_hertz %.1f >= upper_edge_hertz %.1f" % (lower_edge_hertz, upper_edge_hertz))
    if upper_edge_hertz > nyquist_hertz:
        raise ValueError("upper_edge_hertz %.1f is greater than Nyquist %.1f" % (upper_edge_hertz, nyquist_hertz))
    if (num_mel_bins, num_spectrogram_bins, audio_sample_rate, lower_edge_hertz, upper_edge_hertz) == \
            (64, 257, 16000, 125, 7500):
        return _engine.front_end_tables()[1]
    bins_mel = hertz_to_mel(np.linspace(0.0, nyquist_hertz, num_spectrogram_bins))
    edges = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    weights = np.empty((num_spectrogram_bins, num_mel_bins))
    for i in range(num_mel_bins):
        lo, mid, hi = edges[i:i + 3]
        weights[:, i] = np.maximum(0.0, np.minimum((bins_mel - lo) / (mid - lo), (hi - bins_mel) / (hi - mid)))
    weights[0, :] = 0.0
    return weights


def _as_cuda_wave(data):
    if isinstance(data, torch.Tensor):
        t = data
    else:
        t = torch.from_numpy(np.ascontiguousarray(data))
    if t.dim() != 1:
        raise ValueError("expected a 1-D waveform")
    return t.to(device="cuda", dtype=torch.float32)


def _check_vggish_params(audio_sample_rate, log_offset, window_length_secs, hop_length_secs, kwargs):
    got = (audio_sample_rate, log_offset, int(round(audio_sample_rate * window_length_secs)),
           int(round(audio_sample_rate * hop_length_secs)), kwargs.get("num_mel_bins", 20),
           kwargs.get("lower_edge_hertz", 125.0), kwargs.get("upper_edge_hertz", 3800.0))
    if got != (16000, 0.01, 400, 160, 64, 125, 7500):
        raise NotImplementedError(
            "the CUDA front end is specialised for the VGGish parameters (16 kHz, 25 ms window, 10 ms hop, 64 mel "
            "bands 125-7500 Hz, log offset 0.01); got %r" % (got,))


def stft_magnitude(signal, fft_length, hop_length=None, window_length=None):
    """(num_frames, fft_length / 2 + 1) magnitudes of the Hann-windowed frames (reference mel_features.py:71-92),
    computed in float64 on the GPU (vmb_stft_magnitude: the fused log-mel kernel never materialises the magnitudes, so
    the standalone function has its own kernel over all 257 bins).  numpy in -> float64 numpy out like the reference;
    CUDA tensor in -> float64 CUDA tensor.  Specialised for the VGGish framing (fft 512, hop 160, window 400)."""
    if (int(fft_length), hop_length, window_length) != (512, 160, 400):
        raise NotImplementedError("the CUDA front end is specialised for the VGGish framing (fft_length 512, "
                                  "hop_length 160, window_length 400); got %r" % ((fft_length, hop_length, window_length),))
    if isinstance(signal, torch.Tensor):
        if signal.dim() != 1:
            raise ValueError("expected a 1-D waveform")
        return _engine.stft_magnitude(signal.to(device="cuda", dtype=torch.float64))
    sig = np.ascontiguousarray(signal, dtype=np.float64)
    if sig.ndim != 1:
        raise ValueError("expected a 1-D waveform")
    return _engine.stft_magnitude(torch.from_numpy(sig).to("cuda")).cpu().numpy()


def log_mel_spectrogram(data, audio_sample_rate=8000, log_offset=0.0, window_length_secs=0.025,
                        hop_length_secs=0.010, **kwargs):
    """(num_frames, num_mel_bins) log-mel spectrogram of a mono waveform (reference mel_features.py:192-223),
    computed on the GPU.  numpy in -> float64 numpy out (like the reference); CUDA tensor in -> fp32 CUDA tensor."""
    _check_vggish_params(audio_sample_rate, log_offset, window_length_secs, hop_length_secs, kwargs)
    out = _engine.logmel(_as_cuda_wave(data))[0]
    if isinstance(data, torch.Tensor):
        return out
    return out.cpu().numpy().astype(np.float64)
