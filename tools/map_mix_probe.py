"""Which layers of the VGGish body decide the ranking metric?  (GPU box; investigation tool, not product.)

Runs the body layer by layer through the C-ABI layer entry points in fp16 (or bf16) up to / from a cut layer and the rest
in fp32 with torch on the device (a stand-in for the split-precision kernels), then the library head, and prints the mAP
of the scores against the fp32 CPU oracle on the fixed label sets of tests/test_gpu_parity.py plus 20 more label seeds.

    python tools/map_mix_probe.py
"""
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

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = torch.device("cuda:0")
CONV = ((3, 48, 32, 64, 128, 1), (6, 24, 16, 128, 256, 0), (8, 24, 16, 256, 256, 1), (11, 12, 8, 256, 512, 0),
        (13, 12, 8, 512, 512, 1))
FC = ((0, 12288, 4096), (2, 4096, 4096), (4, 4096, 128))
NAMES = ("conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "fc1", "fc2", "fc3")


def body(sd, x, lowp, dtype=1):
    """x (n, 96, 64) fp32 CUDA.  lowp: set of layer indices 0..8 run by the library in 16 bits; the others in fp32 torch.
    Activations travel as fp32 NCHW / (n, features) between layers and are cast where a 16-bit layer consumes them."""
    L, st = _lib.lib(), engine.stream_ptr
    tdt = torch.float16 if dtype else torch.bfloat16
    n = x.shape[0]
    a = x[:, None]                                            # NCHW fp32
    w0, b0 = sd["features.0.weight"].to(DEV), sd["features.0.bias"].to(DEV)
    if 0 in lowp:
        o = torch.empty(n, 48, 32, 64, device=DEV, dtype=tdt)
        engine.check(L.vmb_conv1_relu_pool_ex(x.contiguous().data_ptr(), w0.contiguous().data_ptr(), b0.data_ptr(),
                                              o.data_ptr(), n, dtype, st()), "conv1")
        a = o.permute(0, 3, 1, 2).float()
    else:
        a = F.max_pool2d(F.relu(F.conv2d(a, w0, b0, padding=1)), 2, 2)
    for i, (key, H, W, cin, cout, pool) in enumerate(CONV, start=1):
        w, b = sd[f"features.{key}.weight"].to(DEV), sd[f"features.{key}.bias"].to(DEV)
        if i in lowp:
            xin = a.permute(0, 2, 3, 1).contiguous().to(tdt)
            wk = w.permute(0, 2, 3, 1).contiguous().reshape(cout, 9 * cin).to(tdt)
            o = torch.empty((n, H // 2, W // 2, cout) if pool else (n, H, W, cout), device=DEV, dtype=tdt)
            engine.check(L.vmb_conv3x3_relu_ex(xin.data_ptr(), wk.data_ptr(), b.data_ptr(), o.data_ptr(), n, H, W, cin,
                                               cout, pool, dtype, st()), "conv")
            a = o.permute(0, 3, 1, 2).float()
        else:
            a = F.relu(F.conv2d(a, w, b, padding=1))
            if pool:
                a = F.max_pool2d(a, 2, 2)
    a = a.permute(0, 2, 3, 1).contiguous().reshape(n, 12288)
    for j, (key, fin, fout) in enumerate(FC):
        w, b = sd[f"embeddings.{key}.weight"].to(DEV), sd[f"embeddings.{key}.bias"].to(DEV)
        if 6 + j in lowp:
            o = torch.empty(n, fout, device=DEV, dtype=torch.float32)
            engine.check(L.vmb_linear_ex(a.to(tdt).contiguous().data_ptr(), w.to(tdt).contiguous().data_ptr(), b.data_ptr(),
                                         o.data_ptr(), 1, 1, n, fout, fin, dtype, st()), "fc")
            a = o if j == 2 else o.to(tdt).float()           # the 16-bit kernels round fc1 / fc2 outputs to 16 bits
        else:
            a = F.relu(F.linear(a, w, b))
    return a


def main():
    n = 128
    vsd = synth.vggish_state_dict(0)
    hsd = synth.mla_state_dict((2, 1), 128, 600, 527, 10, seed=2)
    waves = synth.make_clips(100, n)
    ex = np.concatenate([frontend_np.waveform_to_examples(w.astype(np.float64)) for w in waves]).astype(np.float32)
    with torch.no_grad():
        emb_ref = model_torch.vgg_forward(vsd, torch.from_numpy(ex)[:, None])
        want = model_torch.mla_forward(hsd, emb_ref.reshape(n, 10, 128), (2, 1)).numpy()
    head = engine.MlaHandle(hsd, (2, 1), 128, 600, 527, 10, DEV)
    x = torch.from_numpy(ex).to(DEV)
    label_sets = {"test(p=.2,seed=3)": synth.multihot_labels(n, 527, p=0.2, seed=3)}
    for s in range(20):
        label_sets[f"p=.2 seed={10 + s}"] = synth.multihot_labels(n, 527, p=0.2, seed=10 + s)
    for s in range(10):
        label_sets[f"p=.05 seed={s}"] = synth.multihot_labels(n, 527, p=0.05, seed=s)
    ranked = (want >= np.quantile(want, 0.8, axis=0, keepdims=True)).astype(np.int64)
    ref_map = {k: synth.mean_average_precision(v, want) for k, v in label_sets.items()}
    ref_ranked = synth.mean_average_precision(ranked, want)
    configs = {"all fp32 (torch)": set(), "all fp16": set(range(9)), "all bf16": set(range(9))}
    for k in range(1, 9):
        configs[f"fp16 up to {NAMES[k - 1]}, fp32 from {NAMES[k]}"] = set(range(k))
    for k in range(1, 9):
        configs[f"fp32 up to {NAMES[k - 1]}, fp16 from {NAMES[k]}"] = set(range(k, 9))
    out = {}
    for name, lowp in configs.items():
        with torch.no_grad():
            emb = body(vsd, x, lowp, dtype=0 if "bf16" in name else 1)
            scores = head.forward(emb.reshape(n, 10, 128)).cpu().numpy()
        d = {k: synth.mean_average_precision(v, scores) - ref_map[k] for k, v in label_sets.items()}
        vals = np.array(list(d.values()))
        mism = sum(f"{ref_map[k] + d[k]:.3f}" != f"{ref_map[k]:.3f}" for k in d)
        r = {"emb_rel_max": float((emb.cpu() - emb_ref).abs().max() / emb_ref.abs().max()),
             "scores_max_abs": float(np.abs(scores - want).max()), "dmap_test_set": d["test(p=.2,seed=3)"],
             "dmap_abs_mean": float(np.abs(vals).mean()), "dmap_abs_max": float(np.abs(vals).max()),
             "three_decimal_mismatches": f"{mism}/{len(d)}",
             "dmap_ranked": synth.mean_average_precision(ranked, scores) - ref_ranked}
        out[name] = r
        print(f"{name:40s} emb {r['emb_rel_max']:.2e} scores {r['scores_max_abs']:.2e} dmAP test {r['dmap_test_set']:+.2e} "
              f"|dmAP| mean {r['dmap_abs_mean']:.2e} max {r['dmap_abs_max']:.2e} 3-dec mismatches {r['three_decimal_mismatches']} "
              f"ranked {r['dmap_ranked']:+.2e}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_map_mix_probe.json"), "w") as fh:
        json.dump({"oracle_map": ref_map, "oracle_ranked": ref_ranked, "configs": out}, fh, indent=1)


if __name__ == "__main__":
    main()
