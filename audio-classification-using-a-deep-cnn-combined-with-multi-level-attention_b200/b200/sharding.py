"""Partitioning of the work across the GPUs of one box.

Clips are independent in eval mode — nothing in model.py:217-269 or vggish.py:21-31 mixes rows of a batch — so
the inference path shards purely by batch: contiguous slices of the clip axis, weights replicated, NO collective.
A long stream shards by example: example j covers frames 96j .. 96j+95 (vggish_input.py:73-76) and frame i starts
at sample 160 i (mel_features.py:42-45), so a slice of examples [e0, e1) needs samples
[15360 e0, 15360 (e1-1) + 15600): cuts at multiples of 96*160 = 15 360 samples plus a 240-sample tail keep every
frame index identical to the unsharded run.
"""
from __future__ import annotations

import os
from typing import List, Tuple

EXAMPLE_HOP_SAMPLES = 96 * 160          # 15 360
EXAMPLE_SPAN_SAMPLES = 95 * 160 + 400   # 15 600: samples one example reads


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of n_items for `rank` of `world` (first ranks take the remainder)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_items < 0:
        raise ValueError("negative item count")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def num_frames(n_samples: int) -> int:
    """mel_features.py:42 with window 400 / hop 160 (floor division, may be <= 0)."""
    return 1 + (n_samples - 400) // 160


def num_examples(n_samples: int) -> int:
    """Examples waveform_to_examples yields (vggish_input.py:66-76); raises like the reference below 400 samples."""
    nf = num_frames(n_samples)
    if nf < 0:
        raise ValueError("negative dimensions are not allowed")
    return 0 if nf < 96 else 1 + (nf - 96) // 96


def stream_sample_range(e0: int, e1: int) -> Tuple[int, int]:
    """Samples [s0, s1) that examples [e0, e1) of a stream read."""
    if e1 <= e0:
        return e0 * EXAMPLE_HOP_SAMPLES, e0 * EXAMPLE_HOP_SAMPLES
    return e0 * EXAMPLE_HOP_SAMPLES, (e1 - 1) * EXAMPLE_HOP_SAMPLES + EXAMPLE_SPAN_SAMPLES


def stream_chunks(n_samples: int, examples_per_chunk: int, rank: int = 0,
                  world: int = 1) -> List[Tuple[int, int, int, int]]:
    """[(example_start, example_stop, sample_start, sample_stop), ...] covering this rank's share of the stream in
    chunks of at most `examples_per_chunk` examples."""
    if examples_per_chunk < 1:
        raise ValueError("examples_per_chunk must be positive")
    e_lo, e_hi = shard_bounds(num_examples(n_samples), rank, world)
    out = []
    for e0 in range(e_lo, e_hi, examples_per_chunk):
        e1 = min(e_hi, e0 + examples_per_chunk)
        out.append((e0, e1) + stream_sample_range(e0, e1))
    return out


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU); (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))
