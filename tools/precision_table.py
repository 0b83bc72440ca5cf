"""Side-by-side precision table of the three modes of the VGGish body (bf16 / fp16 / split) for DESIGN §3.

Runs on the GPU box.  For 128 seeded clips (the clips of tests/test_gpu_parity.py::test_map_identical_to_three_decimals)
it prints, per mode: per-layer rel-max error and cosine against the fp32 CPU oracle (bf16 and fp16: the layer-level
C-ABI calls chained by hand, 20 examples), the error of the embeddings and of the scores, the uint8 LSB histogram after
the PCA postprocessor, and the macro mAP on the two fixed label sets next to the oracle's.  The oracle (oracle/) is the
checker here, as in the tests.

    python tools/precision_table.py [--clips 128] [--out gpurun_out/precision_table.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-classification-using-a-deep-cnn-combined-with-multi-level-attention_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from b200 import _lib, engine, synth  # noqa: E402
from oracle import frontend_np, model_torch  # noqa: E402

DEV = torch.device("cuda:0")
CONV = ((3, 48, 32, 64, 128, 1), (6, 24, 16, 128, 256, 0), (8, 24, 16, 256, 256, 1), (11, 12, 8, 256, 512, 0),
        (13, 12, 8, 512, 512, 1))
FC = ((0, 12288, 4096), (2, 4096, 4096), (4, 4096, 128))
NAMES = ("conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "fc1", "fc2", "fc3")


def err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return {"rel_max": (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-12),
            "cos": F.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0).item()}


def layer_chain(sd, x, dtype):
    """The body layer by layer through the C-ABI layer entry points; returns the activations in NCHW / (n, features)."""
    L, st = _lib.lib(), engine.stream_ptr
    tdt = torch.float16 if dtype else torch.bfloat16
    n = x.shape[0]
    acts = []
    w0 = sd["features.0.weight"].to(DEV).contiguous()
    b0 = sd["features.0.bias"].to(DEV)
    a = torch.empty(n, 48, 32, 64, device=DEV, dtype=tdt)
    engine.check(L.vmb_conv1_relu_pool_ex(x.data_ptr(), w0.data_ptr(), b0.data_ptr(), a.data_ptr(), n, dtype, st()), "conv1")
    acts.append(a.permute(0, 3, 1, 2))
    for key, H, W, cin, cout, pool in CONV:
        w = sd[f"features.{key}.weight"].to(DEV).permute(0, 2, 3, 1).contiguous().reshape(cout, 9 * cin).to(tdt)
        b = sd[f"features.{key}.bias"].to(DEV)
        o = torch.empty((n, H // 2, W // 2, cout) if pool else (n, H, W, cout), device=DEV, dtype=tdt)
        engine.check(L.vmb_conv3x3_relu_ex(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), n, H, W, cin, cout,
                                           pool, dtype, st()), "conv")
        a = o
        acts.append(a.permute(0, 3, 1, 2))
    a = a.reshape(n, 12288)
    for i, (key, fin, fout) in enumerate(FC):
        w = sd[f"embeddings.{key}.weight"].to(DEV).to(tdt).contiguous()
        b = sd[f"embeddings.{key}.bias"].to(DEV)
        f32 = 1 if i == 2 else 0
        o = torch.empty(n, fout, device=DEV, dtype=torch.float32 if f32 else tdt)
        engine.check(L.vmb_linear_ex(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), f32, 1, n, fout, fin, dtype,
                                     st()), "fc")
        a = o
        acts.append(a)
    torch.cuda.synchronize()
    return acts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "precision_table.json"))
    args = ap.parse_args()
    n = args.clips
    vsd = synth.vggish_state_dict(0)
    hsd = synth.mla_state_dict((2, 1), 128, 600, 527, 10, seed=2)
    waves = synth.make_clips(100, n)
    ex = np.concatenate([frontend_np.waveform_to_examples(w.astype(np.float64)) for w in waves]).astype(np.float32)
    with torch.no_grad():
        emb_ref = model_torch.vgg_forward(vsd, torch.from_numpy(ex)[:, None])
        want = model_torch.mla_forward(hsd, emb_ref.reshape(n, 10, 128), (2, 1)).numpy()
        per_layer_ref = []
        model_torch.vgg_forward(vsd, torch.from_numpy(ex[:20])[:, None], per_layer_ref)
    eig, means = synth.pca_params(1)
    q_ref = model_torch.postprocess(eig, means, emb_ref).numpy()
    # label sets: SURVEY 8d's (Bernoulli(0.05) multi-hot over the whole batch, every class >= 1 positive, seed 3 = the
    # defaults of synth.multihot_labels), round 1's (first 128 clips, p = 0.2, seed 3), labels that follow the oracle's
    # own ranking (top 20 % per class), and 20 more seeds of the SURVEY recipe for the spread
    sets = {"survey_8d": (slice(0, n), synth.multihot_labels(n, 527)),
            "round1_p02_128": (slice(0, 128), synth.multihot_labels(128, 527, p=0.2, seed=3)),
            "oracle_ranked": (slice(0, n), (want >= np.quantile(want, 0.8, axis=0, keepdims=True)).astype(np.int64))}
    for sd_ in range(20):
        sets[f"seed{100 + sd_}"] = (slice(0, n), synth.multihot_labels(n, 527, seed=100 + sd_))

    def maps(scores):
        return {k: synth.mean_average_precision(lab, scores[sl]) for k, (sl, lab) in sets.items()}

    ref_maps = maps(want)
    out = {"clips": n, "oracle": {"mAP": ref_maps}}
    head = engine.MlaHandle(hsd, (2, 1), 128, 600, 527, 10, DEV)
    wave_dev = torch.from_numpy(waves).to(DEV)
    x20 = torch.from_numpy(ex[:20]).to(DEV)
    for mode in ("bf16", "fp16", "split"):
        h = engine.VggishHandle(vsd, DEV, precision=mode)
        scores, emb = engine.Pipeline(h, head).forward(wave_dev, want_embeddings=True)
        h.check_saturation()
        got = scores.cpu().numpy()
        q = engine.postprocess(emb, eig.to(DEV), means.to(DEV)).cpu().numpy()
        d = np.abs(q - q_ref).astype(np.int64)
        r = {"embeddings": err(emb, emb_ref), "scores_max_abs": float(np.abs(got - want).max()),
             "uint8_lsb_histogram": np.bincount(d.ravel()).tolist(),
             "uint8_exact_frac": float((d == 0).mean()), "uint8_within_1_frac": float((d <= 1).mean()),
             "mAP": maps(got)}
        r["mAP_three_decimals_equal"] = {k: f"{v:.3f}" == f"{ref_maps[k]:.3f}" for k, v in r["mAP"].items()}
        d = np.array([r["mAP"][k] - ref_maps[k] for k in ref_maps if k.startswith("seed")])
        r["mAP_abs_delta_over_20_seeds"] = {"mean": float(np.abs(d).mean()), "max": float(np.abs(d).max()),
                                            "three_decimals_equal": int(sum(r["mAP_three_decimals_equal"][k]
                                                                            for k in ref_maps if k.startswith("seed")))}
        if mode != "split":
            acts = layer_chain(vsd, x20, 1 if mode == "fp16" else 0)
            r["layers"] = {nm: err(a, ref) for nm, a, ref in zip(NAMES, acts, per_layer_ref)}
            r["max_activation"] = {nm: float(ref.abs().max()) for nm, ref in zip(NAMES, per_layer_ref)}
        out[mode] = r
        h.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    o = out["oracle"]["mAP"]
    main3 = ("survey_8d", "round1_p02_128", "oracle_ranked")
    print("oracle: mAP " + "  ".join(f"{k} {o[k]:.5f}" for k in main3))
    for mode in ("bf16", "fp16", "split"):
        r = out[mode]
        print(f"{mode:5s}: emb rel-max {r['embeddings']['rel_max']:.2e} cos {r['embeddings']['cos']:.7f}  scores "
              f"{r['scores_max_abs']:.2e}  mAP " + "  ".join(f"{k} {r['mAP'][k]:.5f}" for k in main3) +
              f"  |dmAP| over 20 seeds mean {r['mAP_abs_delta_over_20_seeds']['mean']:.2e} max "
              f"{r['mAP_abs_delta_over_20_seeds']['max']:.2e} 3-dec equal "
              f"{r['mAP_abs_delta_over_20_seeds']['three_decimals_equal']}/20  uint8 exact {r['uint8_exact_frac']:.4f} <=1 "
              f"{r['uint8_within_1_frac']:.4f} hist {r['uint8_lsb_histogram'][:8]}")
        for nm, e in r.get("layers", {}).items():
            print(f"        {nm:8s} rel-max {e['rel_max']:.2e} cos {e['cos']:.7f}")


if __name__ == "__main__":
    main()
