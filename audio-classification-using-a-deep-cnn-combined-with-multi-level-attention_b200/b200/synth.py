"""Seeded synthetic inputs and weights: no dataset or checkpoint is reachable offline
(vggish.py:147,155 download by URL), so every parity test and benchmark uses these.

Weights use He-normal init so activations keep a non-trivial scale through all layers (with PyTorch's default init
the embeddings become bias-dominated and any kernel "passes").  State-dict key names and
shapes are the reference's.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SAMPLE_RATE = 16000
CLIP_SAMPLES = 160000  # 10 s

VGG_CONV = ((0, 1, 64), (3, 64, 128), (6, 128, 256), (8, 256, 256), (11, 256, 512), (13, 512, 512))
VGG_FC = ((0, 12288, 4096), (2, 4096, 4096), (4, 4096, 128))


def make_clip(index: int, n_samples: int = CLIP_SAMPLES) -> np.ndarray:
    """Clip `index`: float64 waveform from default_rng(index); four signal families rotate by index % 4."""
    rng = np.random.default_rng(index)
    t = np.arange(n_samples) / SAMPLE_RATE
    fam = index % 4
    if fam == 0:
        x = rng.uniform(-1.0, 1.0, n_samples)
    elif fam == 1:
        x = 0.5 * np.sin(2 * np.pi * (200.0 + 300.0 * t) * t) + 1e-2 * rng.standard_normal(n_samples)
    elif fam == 2:
        x = rng.normal(0.0, 0.1, n_samples) * (1.0 + np.sin(2 * np.pi * 2.0 * t))
    else:
        x = np.round(rng.normal(0.0, 0.05, n_samples) * 32768.0) / 32768.0
    return x


def make_clips(first: int, count: int, n_samples: int = CLIP_SAMPLES, dtype=np.float32) -> np.ndarray:
    return np.stack([make_clip(first + i, n_samples) for i in range(count)]).astype(dtype)


def fast_clips(first: int, count: int, n_samples: int = CLIP_SAMPLES) -> torch.Tensor:
    """Cheap large-batch generator for benchmarks (same four families, torch RNG; fp32 CPU tensor)."""
    g = torch.Generator().manual_seed(1234 + first)
    t = torch.arange(n_samples, dtype=torch.float64) / SAMPLE_RATE
    out = torch.empty(count, n_samples, dtype=torch.float32)
    chirp = (0.5 * torch.sin(2 * math.pi * (200.0 + 300.0 * t) * t)).float()
    am = (1.0 + torch.sin(2 * math.pi * 2.0 * t)).float()
    for i in range(count):
        fam = (first + i) % 4
        if fam == 0:
            out[i] = torch.rand(n_samples, generator=g) * 2 - 1
        elif fam == 1:
            out[i] = chirp + 1e-2 * torch.randn(n_samples, generator=g)
        elif fam == 2:
            out[i] = 0.1 * torch.randn(n_samples, generator=g) * am
        else:
            out[i] = torch.round(0.05 * torch.randn(n_samples, generator=g) * 32768.0) / 32768.0
    return out


def _he(gen: torch.Generator, *shape: int, fan_in: int) -> torch.Tensor:
    return torch.randn(*shape, generator=gen) * math.sqrt(2.0 / fan_in)


def vggish_state_dict(seed: int = 0) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for idx, cin, cout in VGG_CONV:
        sd[f"features.{idx}.weight"] = _he(g, cout, cin, 3, 3, fan_in=cin * 9)
        sd[f"features.{idx}.bias"] = torch.randn(cout, generator=g) * 0.05
    for idx, fin, fout in VGG_FC:
        sd[f"embeddings.{idx}.weight"] = _he(g, fout, fin, fan_in=fin)
        sd[f"embeddings.{idx}.bias"] = torch.randn(fout, generator=g) * 0.05
    return sd


def pca_params(seed: int = 1, gain: float = 0.35):
    """Synthetic PCA: seeded orthogonal matrix x a bounded diagonal gain in [gain/2, gain] (SURVEY §7 H2 asks
    for a stated, bounded gain), means = seeded.  Returns (eigen_vectors (128,128), means (128,1))."""
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(128, 128, generator=g, dtype=torch.float64))
    d = gain * (0.5 + 0.5 * torch.rand(128, generator=g, dtype=torch.float64))
    eig = (d[:, None] * q).float()
    means = (torch.rand(128, 1, generator=g) * 1.5)
    return eig, means


def _bn(g: torch.Generator, n: int, prefix: str, sd: dict) -> None:
    sd[prefix + ".weight"] = 0.5 + torch.rand(n, generator=g)
    sd[prefix + ".bias"] = torch.randn(n, generator=g) * 0.1
    sd[prefix + ".running_mean"] = torch.randn(n, generator=g) * 0.2
    sd[prefix + ".running_var"] = 0.8 + 0.8 * torch.rand(n, generator=g)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def mla_state_dict(model_conf=(2, 1), emb_in: int = 128, hidden: int = 600, n_classes: int = 527, t_steps: int = 10,
                   seed: int = 2, include_fcf: bool = True) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for lvl, n_fc in enumerate(model_conf):
        p = f"embedded_mappings.{lvl}"
        _bn(g, t_steps, p + ".norm0", sd)
        if lvl == 0:
            # the embeddings entering level 0 are post-ReLU VGGish outputs with a large scale; give norm0
            # statistics of that order so the head is not saturated
            sd[p + ".norm0.running_mean"] = 1.0 + torch.rand(t_steps, generator=g)
            sd[p + ".norm0.running_var"] = 4.0 + 4.0 * torch.rand(t_steps, generator=g)
        for j in range(n_fc):
            fin = emb_in if (lvl == 0 and j == 0) else hidden
            sd[f"{p}.fc.{j}.weight"] = _he(g, hidden, fin, fan_in=fin)
            sd[f"{p}.fc.{j}.bias"] = torch.randn(hidden, generator=g) * 0.05
        for j in range(n_fc):
            _bn(g, t_steps, f"{p}.norms.{j}", sd)
    for lvl in range(len(model_conf)):
        p = f"attention_modules.{lvl}"
        sd[p + ".fcv.weight"] = _he(g, n_classes, hidden, fan_in=hidden)
        sd[p + ".fcv.bias"] = torch.randn(n_classes, generator=g) * 0.05
        if include_fcf:
            sd[p + ".fcf.weight"] = _he(g, n_classes, hidden, fan_in=hidden)
            sd[p + ".fcf.bias"] = torch.randn(n_classes, generator=g) * 0.05
        _bn(g, t_steps, p + ".normv", sd)
        _bn(g, t_steps, p + ".normf", sd)
    L = len(model_conf)
    sd["fc.weight"] = _he(g, n_classes, L * n_classes, fan_in=L * n_classes)
    sd["fc.bias"] = torch.randn(n_classes, generator=g) * 0.05
    _bn(g, n_classes, "norm", sd)
    return sd


def multihot_labels(batch: int, n_classes: int = 527, p: float = 0.05, seed: int = 3) -> np.ndarray:
    """Seeded Bernoulli(p) multi-hot labels with every class forced to have at least one positive."""
    rng = np.random.default_rng(seed)
    y = (rng.random((batch, n_classes)) < p).astype(np.int64)
    for k in np.nonzero(y.sum(axis=0) == 0)[0]:
        y[rng.integers(0, batch), k] = 1
    return y


def mean_average_precision(labels: np.ndarray, scores: np.ndarray) -> float:
    """Macro mAP (sklearn.metrics.average_precision_score semantics, ties handled like sklearn: by thresholds)."""
    from sklearn.metrics import average_precision_score
    return float(average_precision_score(labels, scores, average="macro"))
