"""Drop-in for torchvggish/vggish_input.py: waveform -> VGGish examples, computed by the fused CUDA front end."""
import numpy as np
import torch

from b200 import engine as _engine

from . import vggish_params


def waveform_to_examples(data, sample_rate, return_tensor=True):
    """Converts an audio waveform into VGGish examples (reference vggish_input.py:30-82).

    data: np.ndarray (or torch tensor) of shape (samples,) or (samples, channels) — the reference averages
      axis 1 (vggish_input.py:49-50) — nominally in [-1, 1].
    sample_rate: must be 16000; the reference resamples other rates with the third-party `resampy`, which is
      outside this build -> NotImplementedError.
    Returns (num_examples, 1, 96, 64) float32 CUDA tensor, or with return_tensor=False the reference's
      (num_examples, 96, 64) float64 numpy array.  Fewer than 400 samples raises ValueError like numpy does in
      the reference; fewer than 15 600 samples gives zero examples.
    """
    if isinstance(data, torch.Tensor):
        wave = data.to(device="cuda")
        if wave.dim() > 1:
            wave = wave.double().mean(dim=1)
    else:
        data = np.asarray(data)
        if len(data.shape) > 1:
            data = np.mean(data, axis=1)
        # the kernel reads fp32 samples: round on the host (the same round-to-nearest the device cast would do) so that
        # only half of a float64 waveform's bytes cross the host -> device link
        wave = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(device="cuda")
    if sample_rate != vggish_params.SAMPLE_RATE:
        raise NotImplementedError("resampling to 16 kHz (resampy in the reference) is outside the B200 path")
    examples = _engine.examples_from_wave(wave.to(torch.float32))
    if return_tensor:
        return examples[:, None, :, :]
    return examples.cpu().numpy().astype(np.float64)


def wavfile_to_examples(wav_file, return_tensor=True):
    """16-bit PCM WAV -> examples (reference vggish_input.py:85-99).  Uses the standard-library `wave` reader
    instead of the third-party `soundfile`; samples are scaled by 1/32768 like the reference."""
    import wave as _wave
    with _wave.open(wav_file, "rb") as wf:
        assert wf.getsampwidth() == 2, "Bad sample type: %r" % (wf.getsampwidth(),)
        sr = wf.getframerate()
        pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype="<i2")
        if wf.getnchannels() > 1:
            pcm = pcm.reshape(-1, wf.getnchannels())
    if pcm.ndim == 1 and sr == vggish_params.SAMPLE_RATE:
        # mono 16 kHz: the int16 samples go to the device as they are and are scaled by 1/32768 there (exact in fp32,
        # bit-identical to the float path); half the host-to-device bytes
        dev_pcm = torch.from_numpy(np.ascontiguousarray(pcm)).to(device="cuda")
        per = _engine.num_examples(dev_pcm.shape[0])
        if per < 0:
            raise ValueError("negative dimensions are not allowed")
        examples = _engine.logmel_pcm16(dev_pcm, per * 96).view(per, 96, 64)
        if return_tensor:
            return examples[:, None, :, :]
        return examples.cpu().numpy().astype(np.float64)
    samples = pcm / 32768.0
    return waveform_to_examples(samples, sr, return_tensor)
