"""CPU emulation of the log-mel kernels' arithmetic (no GPU needed): operand splitting into bf16 / fp16 planes, the
product order, and an fp32 accumulator that TRUNCATES after every K = 16 MMA (what tcgen05.mma does), against the
float64 reference (oracle/frontend_np.py).  This is how the operand format and product order of
csrc/logmel_tc.cu: logmel_eo_kernel were chosen before any GPU time was spent; the predictions for the shipped scheme
(fp16 hi + lo, scaled, three products, centred even/odd DFT) were 3e-5 on the chirp family and 1.7e-4 on the full-scale
tone, the GPU measured 2.1e-5 and 1.7e-4.

    python tools/logmel_precision_emulation.py            # ~3 minutes
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200"))
from b200 import synth  # noqa: E402
from oracle import frontend_np  # noqa: E402

HANN = frontend_np.periodic_hann(400)
MEL = frontend_np.mel_matrix()
BINS = np.arange(4, 244)


def bf16(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy().astype(np.float64)


def fp16(x):
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def r11(x):
    """11 significant bits, unlimited exponent: fp16 planes after the power-of-two scaling (no subnormals)."""
    m, e = np.frexp(np.asarray(x, dtype=np.float64))
    return np.ldexp(np.round(m * 2048) / 2048, e)


def split(x, rnd, planes):
    out, r = [], np.asarray(x, dtype=np.float64)
    for _ in range(planes):
        p = rnd(r)
        out.append(p)
        r = (r - p).astype(np.float32).astype(np.float64)
    return out


def trunc32(x):
    f = np.asarray(x, dtype=np.float64).astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float64)


def gemm(A, B, rnd, planes, products, passes=None):
    """passes: list of product lists accumulated one after the other over all of K (None: all interleaved per K block)."""
    Ap, Bp = split(A, rnd, planes), split(B, rnd, planes)
    acc = np.zeros((A.shape[0], B.shape[1]))
    K = A.shape[1]
    for prods in (passes or [products]):
        for k0 in range(0, K, 32):
            for i, j in prods:
                for k1 in range(k0, min(K, k0 + 32), 16):
                    acc = trunc32(acc + Ap[i][:, k1:k1 + 16] @ Bp[j][k1:k1 + 16])
    return acc


def frames_of(x):
    n = 1 + (len(x) - 400) // 160
    return x[np.arange(400)[None, :] + 160 * np.arange(n)[:, None]]


def finish(re, im):
    mag = np.sqrt(re ** 2 + im ** 2).astype(np.float32).astype(np.float64)
    return np.log(mag @ MEL[4:244] + 0.01)


def straight(x, **kw):
    fr = frames_of(x.astype(np.float32).astype(np.float64))
    ang = 2 * np.pi * ((BINS[None, :] * np.arange(400)[:, None]) % 512) / 512
    return finish(gemm(fr, HANN[:, None] * np.cos(ang), **kw), gemm(fr, HANN[:, None] * np.sin(ang), **kw))


def centred(x, **kw):
    fr = frames_of(x.astype(np.float32).astype(np.float64))
    m = np.arange(200)
    xp, xm = fr[:, 200 + m], fr[:, 200 - m]
    E = (xp + xm).astype(np.float32).astype(np.float64)
    O = (xp - xm).astype(np.float32).astype(np.float64)
    ang = 2 * np.pi * ((BINS[None, :] * m[:, None]) % 512) / 512
    C = HANN[200 + m][:, None] * np.cos(ang)
    C[0] *= 0.5
    S = HANN[200 + m][:, None] * np.sin(ang)
    return finish(gemm(E, C, **kw), gemm(O, S, **kw))


SIX = [(2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0)]
THREE = [(1, 0), (0, 1), (0, 0)]
SCHEMES = {
    "straight, bf16 x3, six products (first tensor-core kernel)": lambda x: straight(x, rnd=bf16, planes=3, products=SIX),
    "centred, bf16 x3, six products": lambda x: centred(x, rnd=bf16, planes=3, products=SIX),
    "centred, bf16 x3, small products of all K first": lambda x: centred(x, rnd=bf16, planes=3, products=SIX,
                                                                       passes=[SIX[:5], SIX[5:]]),
    "centred, fp16 x2, three products": lambda x: centred(x, rnd=fp16, planes=2, products=THREE),
    "centred, fp16 x2 scaled, three products (shipped)": lambda x: centred(x, rnd=r11, planes=2, products=THREE),
}


def main():
    w = synth.make_clips(0, 4)
    cases = {f"family {i}": w[i][:8000].astype(np.float64) for i in range(4)}
    cases["full-scale int16 tone"] = (np.sin(np.arange(8000) * 0.05) * 20000).astype(np.int16) / 32768.0
    cases["family 0 at 1e-3"] = cases["family 0"] * 1e-3
    for name, fn in SCHEMES.items():
        errs = []
        for cname, x in cases.items():
            errs.append(f"{cname}: {np.abs(fn(x) - frontend_np.log_mel_spectrogram(x)).max():.2e}")
        print(name + "\n    " + "; ".join(errs), flush=True)


if __name__ == "__main__":
    main()
