#!/usr/bin/env python
"""bench_stream.py — BASELINE.json configs[3]: a 1-hour continuous 16 kHz synthetic stream (57.6 M samples, 230 MB
fp32 in pinned host memory) framed into 3 749 examples of 0.96 s, VGGish embeddings, PCA + 8-bit quantisation.
Times b200.stream.embed_stream end to end (H2D of the stream inside the timed region, uint8 result left on the
device and checksummed) and prints one JSON line; --check also runs the CPU oracle on the first minutes of the
stream and reports the uint8 LSB histogram against it."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402



def _json_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when
    the communicator is created), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved
    original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


_OUT = None


def emit(line):
    (_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_OUT or sys.stdout).flush()

def main():
    global _OUT
    _OUT = _json_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=3600)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--chunk", type=int, default=2048, help="examples per chunk")
    ap.add_argument("--check", type=int, default=0, help="examples to verify against the CPU oracle")
    ap.add_argument("--precision", choices=("bf16", "split"), default="bf16",
                    help="bf16 = throughput mode; split = accuracy mode (hi + lo bf16, uint8 output matches fp32)")
    args = ap.parse_args()
    from b200 import engine, sharding, stream, synth
    dev = torch.device("cuda:0")
    engine.require_b200(dev)
    n = args.seconds * 16000
    g = torch.Generator().manual_seed(7)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    wave = (0.3 * torch.sin(2 * np.pi * (110.0 + 40.0 * torch.sin(2 * np.pi * 0.05 * t)) * t)
            + 0.05 * torch.randn(n, generator=g)).pin_memory()
    sd = synth.vggish_state_dict(0)
    vgg = engine.VggishHandle(sd, dev, precision=args.precision)
    eig, means = synth.pca_params(1)
    n_ex = sharding.num_examples(n)
    for _ in range(2):
        emb, q = stream.embed_stream(vgg, wave, eig, means, args.chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        emb, q = stream.embed_stream(vgg, wave, eig, means, args.chunk)
        checksum = int(q.sum().item())          # forces completion, 8 bytes D2H
    dt = (time.perf_counter() - t0) / args.steps
    line = {"metric": "long-form embedding extraction examples/sec (1-hour stream, PCA/uint8)", "value": n_ex / dt,
            "unit": "examples/s", "n_gpus": 1, "steps": args.steps, "seconds_per_stream": dt,
            "realtime_factor": args.seconds / dt, "n_examples": n_ex, "h2d_bytes": n * 4, "checksum": checksum,
            "config": {"workload": f"{args.seconds} s stream, {n} samples, chunks of {args.chunk} examples", "dtype": "bf16" if args.precision == "bf16" else "split bf16 (hi + lo)"}}
    if args.check:
        from oracle import frontend_np, model_torch
        m = args.check
        s0, s1 = sharding.stream_sample_range(0, m)
        ex = frontend_np.waveform_to_examples(wave[s0:s1].numpy().astype(np.float64)).astype(np.float32)
        with torch.no_grad():
            ref = model_torch.vgg_forward(sd, torch.from_numpy(ex)[:, None])
            refq = model_torch.postprocess(eig, means, ref).numpy()
        d = np.abs(q[:m].cpu().numpy().astype(np.float32) - refq).astype(np.int64)
        line["uint8_lsb_histogram_vs_oracle"] = np.bincount(d.ravel()).tolist()
        line["embedding_rel_max_err"] = float((emb[:m].cpu() - ref).abs().max() / ref.abs().max())
    emit(line)


if __name__ == "__main__":
    main()
