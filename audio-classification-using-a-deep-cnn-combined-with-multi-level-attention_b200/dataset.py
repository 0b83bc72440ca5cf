"""Drop-in for the torchvggish branch of the reference's dataset.py: `create_spec`, `split`,
`overlapping_split`, `contiguous_split`, `normalize` with the reference's signatures (dataset.py:276-380), plus
`clips_to_frames`, the batched device-side path that feeds `Ensemble` directly.

The reference builds, per 4 s UrbanSound8K clip, a (64 mel, 384 time) spectrogram out of the <= 4 VGGish examples
(zero where the clip is shorter) and cuts it into T = 10 overlapping (64, 96) windows, 32 steps apart; `Input` later
merely RESHAPES each (64, 96) window to (96, 64) (model.py:98-99) — kept as is.  The log-mel comes from the fused
CUDA front end, the tiling from `vmb_spec_tiles`.  The librosa / h5py / ResNet parts of the module (dataset files,
`create_mfcc`, `plot_spec`, `load_hdf5`) are outside the B200 path and are not provided.
"""
import numpy as np
import torch

from b200 import _lib, engine as _engine
from b200._lib import check, ptr, stream_ptr
from params import *  # noqa: F401,F403
from torchvggish.vggish_input import waveform_to_examples


def create_spec(audio_array, cnn_type, sr, samples_num, x_size, y_size, use_librosa, overlap):
    """(y_size, 4 * x_size) = (64, 384) spectrogram of a clip (dataset.py:276-326), numpy in -> numpy out."""
    if cnn_type != "vggish":
        raise NotImplementedError("only the torchvggish branch is on the B200 path")
    if use_librosa:
        raise NotImplementedError("the librosa spectrogram branch is third-party code outside the B200 path")
    slots = waveform_to_examples(audio_array, sr, return_tensor=False)
    reshaped = np.zeros((4, slots.shape[1], slots.shape[2]))
    reshaped[:slots.shape[0]] = slots[:4]
    return np.concatenate(np.swapaxes(reshaped, 1, 2)[:4], axis=1)


def overlapping_split(spec, num_frames, frame_length):
    step = (spec.shape[1] - frame_length) // (num_frames - 1)
    return np.array([spec[:, i:i + frame_length] for i in range(0, spec.shape[1], step)][:num_frames])


def contiguous_split(spec, num_frames, frame_length):
    return np.array([spec[:, i:i + frame_length] for i in range(0, spec.shape[1], frame_length)][:num_frames])


def split(spec, num_frames, x_size, y_size, overlap):
    frames = overlapping_split(spec, num_frames, x_size) if overlap else contiguous_split(spec, num_frames, x_size)
    for i in range(frames.shape[0]):
        assert frames[i].shape == (y_size, x_size), "{} should be ({},{}); instead is ({},{})".format(
            i, y_size, x_size, frames[i].shape[0], frames[i].shape[1])
    return frames


def normalize(d, _min, _max):
    return (d - _min) / (_max - _min)


def clips_to_frames(waves: torch.Tensor, num_frames: int = T, overlap: bool = True) -> torch.Tensor:
    """waves (n_clips, n_samples <= 4 s) fp32 CUDA 16 kHz -> (n_clips, num_frames, 1, 64, 96) fp32 CUDA: what
    load_hdf5 stores per clip (dataset.py:236-254), computed for the whole batch on the device."""
    if waves.dim() != 2 or not waves.is_cuda or waves.dtype != torch.float32:
        raise ValueError("waves must be a (n_clips, n_samples) fp32 CUDA tensor")
    n_clips = waves.shape[0]
    per = min(4, _engine.num_examples(waves.shape[1]))
    if per < 0:
        raise ValueError("negative dimensions are not allowed")
    ex = (_engine.logmel(waves, per * 96).contiguous() if per else
          torch.empty((n_clips, 0, 64), device=waves.device))
    out = torch.empty((n_clips, num_frames, 1, 64, 96), device=waves.device, dtype=torch.float32)
    with torch.cuda.device(waves.device):
        check(_lib.lib().vmb_spec_tiles(ptr(ex) if per else None, n_clips, per, num_frames, 1 if overlap else 0,
                                        ptr(out), stream_ptr()), "vmb_spec_tiles")
    return out
