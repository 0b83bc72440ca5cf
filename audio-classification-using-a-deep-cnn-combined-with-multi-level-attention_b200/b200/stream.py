"""Long-form embedding extraction (BASELINE.json configs[3]): a continuous 16 kHz stream -> 0.96 s examples ->
VGGish embeddings -> PCA / clamp / 8-bit quantisation (vggish_input.py:66-76, vggish.py:21-31, :62-102).

The stream is cut at multiples of 15 360 samples (96 frames x 160) with a 240-sample tail so that every STFT frame
keeps its index; chunks are independent, so a stream also shards across GPUs by example with no
collective (sharding.stream_chunks).  Results are identical, bit for bit, however the stream is chunked.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, sharding


def embed_stream(vgg: "engine.VggishHandle", wave: torch.Tensor, pca_eigen: Optional[torch.Tensor] = None,
                 pca_means: Optional[torch.Tensor] = None, examples_per_chunk: int = 2048, rank: int = 0,
                 world: int = 1):
    """wave: 1-D fp32 stream, on the host (pinned for full speed) or already on the device.

    Returns this rank's share of the stream's examples: (embeddings fp32 (n, 128), uint8 (n, 128) or None when no
    PCA parameters are given), both on the device, in stream order."""
    if wave.dim() != 1 or wave.dtype != torch.float32:
        raise ValueError("wave must be a 1-D fp32 tensor")
    dev = vgg.device
    n_samples = wave.shape[0]
    chunks = sharding.stream_chunks(n_samples, examples_per_chunk, rank, world)
    n_local = sum(e1 - e0 for e0, e1, _, _ in chunks)
    emb = torch.empty((n_local, 128), device=dev, dtype=torch.float32)
    q = torch.empty((n_local, 128), device=dev, dtype=torch.uint8) if pca_eigen is not None else None
    if pca_eigen is not None:
        pca_eigen = pca_eigen.to(dev, torch.float32).contiguous()
        pca_means = pca_means.to(dev, torch.float32).reshape(-1).contiguous()
    main = torch.cuda.current_stream(dev)
    # The copy stream, its events and the two staging buffers live with the handle and are reused by every call: a fresh
    # stream per call would strand the caching allocator's blocks in that stream's pool (they cannot serve another
    # stream), so each call would cudaMalloc its staging buffers again and, once memory ran short, stall on cudaFree.
    st = getattr(vgg, "_stream_state", None)
    if st is None:
        st = vgg._stream_state = dict(copy=torch.cuda.Stream(device=dev), staged=[None, None],
                                      ready=[torch.cuda.Event(), torch.cuda.Event()],
                                      freed=[torch.cuda.Event(), torch.cuda.Event()])
    copy_stream, staged, ready, freed = st["copy"], st["staged"], st["ready"], st["freed"]
    max_samples = max((s1 - s0 for _, _, s0, s1 in chunks), default=0)
    if not wave.is_cuda:
        copy_stream.wait_stream(main)            # whatever used the buffers in an earlier call has been enqueued on main
        for b in range(2):
            if staged[b] is None or staged[b].numel() < max_samples:
                staged[b] = torch.empty(max_samples, device=dev, dtype=torch.float32)

    def stage(i):
        e0, e1, s0, s1 = chunks[i]
        if wave.is_cuda:
            return wave[s0:s1]
        b = i & 1
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(freed[b])
            staged[b][:s1 - s0].copy_(wave[s0:s1], non_blocking=True)
            ready[b].record(copy_stream)
        return staged[b][:s1 - s0]

    pos = 0
    nxt = stage(0) if chunks else None
    for i, (e0, e1, s0, s1) in enumerate(chunks):
        cur = nxt
        if not wave.is_cuda:
            main.wait_event(ready[i & 1])
        if i + 1 < len(chunks):
            nxt = stage(i + 1)                       # H2D of the next chunk overlaps this chunk's compute
        ex = engine.examples_from_wave(cur)
        out = vgg.forward(ex)
        emb[pos:pos + (e1 - e0)] = out
        if q is not None:
            q[pos:pos + (e1 - e0)] = engine.postprocess(out, pca_eigen, pca_means, want_u8=True)[1]
        if not wave.is_cuda:
            freed[i & 1].record(main)
        pos += e1 - e0
    return emb, q
