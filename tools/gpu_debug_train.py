"""Bring-up helper (not a test): per-tensor gradient errors of the head training step vs the CPU oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200"))
sys.path.insert(0, ROOT)
from b200 import synth, training  # noqa: E402
from oracle import train_torch  # noqa: E402

DEV = torch.device("cuda:0")
for K, conf, batch in [(10, (1,), 32), (10, (2,), 32), (10, (1, 1), 32), (10, (2, 1), 33), (527, (2, 1), 96)]:
    tr = training.HeadTrainer(conf, 128, 600, K, 10, 128, DEV, dropout_p=0.0)
    sd = synth.mla_state_dict(conf, 128, 600, K, 10, seed=7)
    tr.load_state_dict(sd)
    g = torch.Generator().manual_seed(batch)
    x = torch.randn(batch, 10, 128, generator=g)
    labels = torch.randint(0, K, (batch,), generator=g)
    loss, scores = tr.forward_backward(x, labels, want_scores=True)
    ref_loss, ref_scores, ref_grads = train_torch.head_step(sd, x, labels, conf)
    print(f"== K={K} conf={conf} batch={batch}: loss {loss.item():.6f} vs {ref_loss.item():.6f}; "
          f"scores err {(scores.cpu() - ref_scores).abs().max():.2e}")
    for key, shape, off in tr.p_layout:
        got = tr.view(tr.grads, key).cpu().numpy()
        ref = ref_grads[key].numpy()
        print(f"   {np.abs(got - ref).max():9.2e} abs  {np.abs(ref).max():9.2e} max|ref|  "
              f"{np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30):9.2e} rel   {key}")
    tr.close()
